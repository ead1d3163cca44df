"""pipeline::Pipeline and pool::Pool of the reference (src/pipeline.rs:26-137, src/pool.rs:43-250) for this path.

The reference runs every stage on its own thread and connects them with channels.  Here a stage is a CUDA stream and
the channels are CUDA events (`ae_pipeline_*` in the C ABI): `Sender.send(item)` calls every stage's closure once, in
order, on the calling thread; the closures only QUEUE work (kernels, `upload_async` / `download_async` copies) — while a
closure runs, everything the library launches goes to that stage's stream.  Stage k of item i therefore overlaps stage
k+1 of item i-1 on the GPU, items leave in the order they entered, and `report()` is the line each reference stage
prints once a second (processed, active time, rate, utilisation).

    pool = Pool.make(3, maker=lambda: Block(...), resetter=lambda b: None)     # pinned host + device buffers
    tx, rx = (Pipeline.new("h2d", lambda b: b.upload(), depth=3)
              .add_stage("fft", lambda b: b.transform())
              .add_stage("d2h", lambda b: b.download())
              .finish())
    tx.send(pool.take_or_make()); ...; done = rx.recv()

Rules (include/aether_b200.h): a handle with internal scratch (FFT plan, FIR state, Awgn stream) belongs to ONE stage;
a closure must not synchronise (to_numpy, sync, stats.read) nor switch streams."""
from __future__ import annotations

import ctypes as C
import threading

from . import _lib
from ._lib import call

_STAGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p)


class _State:
    def __init__(self, depth: int):
        h = C.c_void_p()
        call("ae_pipeline_create", depth, C.byref(h))
        self.h = h
        self.depth = depth
        self.values = {}            # token -> the object currently travelling under that token
        self.next_token = 1
        self.callbacks = []         # keep the ctypes trampolines alive
        self.error = None

    def close(self):
        if self.h:
            _lib.lib().ae_pipeline_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipeline:
    """Builder, as in the reference: `Pipeline.new(name, op).add_stage(name, op)....finish() -> (Sender, Receiver)`.
    Every `op` maps the object the previous stage returned to the object handed to the next one (`FnMut(I) -> O`)."""

    def __init__(self, state: _State):
        self._s = state

    @classmethod
    def new(cls, name: str, op, depth: int = 3) -> "Pipeline":
        return cls(_State(depth)).add_stage(name, op)

    def add_stage(self, name: str, op) -> "Pipeline":
        s = self._s

        def tramp(_user, _slot, token):
            try:
                s.values[token] = op(s.values[token])
                return 0
            except BaseException as ex:      # a Python exception must not unwind through the C frames
                s.error = ex
                return 1

        cb = _STAGE_FN(tramp)
        s.callbacks.append(cb)
        call("ae_pipeline_add_stage", s.h, name.encode(), C.cast(cb, C.c_void_p), None)
        return self

    def finish(self) -> tuple["Sender", "Receiver"]:
        return Sender(self._s), Receiver(self._s)


class Sender:
    def __init__(self, state: _State):
        self._s = state

    def send(self, item) -> None:
        s = self._s
        token = s.next_token
        s.next_token += 1
        s.values[token] = item
        st = _lib.lib().ae_pipeline_send(s.h, C.c_void_p(token))
        if st != 0:
            s.values.pop(token, None)
            err, s.error = s.error, None
            if err is not None:
                raise err
            _lib.check(st)


class Receiver:
    def __init__(self, state: _State):
        self._s = state

    def recv(self):
        p = C.c_void_p()
        call("ae_pipeline_recv", self._s.h, C.byref(p))
        return self._s.values.pop(p.value)

    def in_flight(self) -> int:
        return int(_lib.lib().ae_pipeline_in_flight(self._s.h))

    def report(self, reset: bool = False) -> list[dict]:
        n = int(_lib.lib().ae_pipeline_stages(self._s.h))
        st = (_lib.PipeStage * max(n, 1))()
        call("ae_pipeline_report", self._s.h, st, n, int(reset))
        return [dict(name=x.name.decode(), processed=int(x.processed), active_ms=x.active_ms, elapsed_ms=x.elapsed_ms,
                     per_second=x.per_second, utilisation_pct=x.utilisation_pct) for x in st[:n]]

    def report_lines(self) -> list[str]:
        """the text the reference's stage threads print (src/pipeline.rs:93-101)"""
        return ["Stage: %-15s : Processed %d in %3.3fs (%9.2f/s); Utilisation: %3.2f%%"
                % (r["name"], r["processed"], r["elapsed_ms"] / 1e3, r["per_second"], r["utilisation_pct"]) for r in self.report()]


class Elem:
    """Guard for an element taken from a Pool (src/pool.rs:196-236): use it as a context manager or call `release()`;
    the element is reset and returned to its pool.  `.val` is the element (the reference derefs the guard)."""

    def __init__(self, pool: "Pool", val):
        self._pool = pool
        self.val = val
        self._live = True

    def release(self) -> None:
        if self._live:
            self._live = False
            self._pool._give_back(self.val)

    def __enter__(self):
        return self.val

    def __exit__(self, *exc) -> bool:
        self.release()
        return False

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Pool:
    """pool::make(initial_len, maker, resetter) (src/pool.rs:43-70): reusable, expensive-to-make objects — pinned host
    buffers and device vectors here.  `take()` is the bounded use (None when empty), `take_or_make()` grows the pool."""

    def __init__(self, initial_len: int, maker, resetter):
        self._maker, self._resetter = maker, resetter
        self._mu = threading.Lock()
        self._elems = []
        for _ in range(initial_len):
            e = maker()
            resetter(e)
            self._elems.append(e)
        self._cap = len(self._elems)

    @classmethod
    def make(cls, initial_len: int, maker, resetter) -> "Pool":
        return cls(initial_len, maker, resetter)

    def clone(self) -> "Pool":
        return self                      # the reference clones an Arc: the same pool

    def take(self):
        with self._mu:
            if not self._elems:
                return None
            return Elem(self, self._elems.pop())

    def take_or_make(self) -> Elem:
        with self._mu:
            if self._elems:
                return Elem(self, self._elems.pop())
            self._cap += 1
        return Elem(self, self._maker())

    def _give_back(self, val) -> None:
        self._resetter(val)
        with self._mu:
            self._elems.append(val)

    def len(self) -> int:
        with self._mu:
            return len(self._elems)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def cap(self) -> int:
        with self._mu:
            return self._cap
