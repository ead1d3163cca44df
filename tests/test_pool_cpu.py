"""pool::Pool (src/pool.rs:260-329): the reference's own three unit tests and its doctest, restated on the Python mirror
(host-side logic, no GPU)."""
import threading

from aether_primitives_b200 import pool
from aether_primitives_b200.pool import Pool


def test_taking():                      # src/pool.rs:264-298
    p: Pool = pool.make(1, lambda: bytearray(), lambda o: o.clear())
    assert p.len() == 1 and p.cap() == 1
    c1 = p.take()
    assert c1 is not None, "First time checkout failed"
    assert p.len() == 0 and p.cap() == 1
    c1.release()
    assert p.len() == 1 and p.cap() == 1
    c1 = p.take()
    assert c1 is not None, "Second checkout failed, when it should have succeeded"
    c2 = p.take()
    assert c2 is None, "Third checkout succeeded when it should have failed"
    c1.release()
    assert p.len() == 1 and p.cap() == 1


def test_resetting():                   # src/pool.rs:300-311
    p = pool.make(1, lambda: bytearray(), lambda o: o.clear())
    with p.take() as first_elem:
        first_elem.extend(range(50))
        assert len(first_elem) == 50, "Vector should contain elements now"
    with p.take() as again:
        assert len(again) == 0          # the resetter ran when the guard was dropped


def test_taking_or_making():            # src/pool.rs:313-329
    p = pool.make(0, lambda: bytearray(), lambda o: o.clear())
    e1 = p.take_or_make()
    assert p.len() == 0 and p.cap() == 1
    e2 = p.take_or_make()
    assert p.len() == 0 and p.cap() == 2
    e1.release()
    e2.release()
    assert p.len() == 2 and p.cap() == 2


def test_doctest_across_threads():      # src/pool.rs:14-41
    p = pool.make(0, lambda: bytearray(50), lambda o: None)
    q = p.clone()
    assert p.len() == 0, "Pool should be empty"
    box = []
    t = threading.Thread(target=lambda: box.append(q.take_or_make()))
    t.start()
    t.join()
    assert p.len() == 0 and p.cap() == 1, "Pool should own 1 element"
    box.pop().release()                 # drop the guard so the element is returned to the pool
    assert p.len() == 1 and p.cap() == 1
    assert not p.is_empty()
