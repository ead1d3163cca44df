//! The reference crate's API on top of libaether_b200.so (C ABI: include/aether_b200.h).
//!
//! UNCOMPILED: the build image has no Rust toolchain and no network.  tests/test_abi_cpu.py checks this
//! file mechanically: every entry point the header declares is called here, no `unimplemented!` / `todo!`
//! is left, and the `-sys` crate it calls is generated from the header.
//!
//! Mapping (reference file:line -> here):
//!   VecOps                 src/vecops.rs:39-89     impl VecOps for DeviceVec (+ `_dev` methods with device operands)
//!   Fft / Cfft             src/fft.rs:48-77,134    impl Fft for CudaFft (host slices) + `_dev` methods
//!   Scale                  src/fft.rs:6-18         scale_args()
//!   Fir                    src/fir.rs:3-22         Fir (the filter method the crate never wrote)
//!   interpolate/downsample src/sampling.rs:7,28,49 sampling::{interpolate, downsample, downsample_sb} (+ `_dev`)
//!   Modulation             src/modulation.rs:94    impl Modulation for Bpsk / Qpsk (host slices) + `_dev` methods
//!   Awgn                   src/noise.rs:9-70       noise::{generator, new, Awgn}
//!   expand / generate      src/sequence.rs:18,47   sequence::{expand, generate_taps}
//!   examples/modem.rs:15   chain::modem_fused;  benches correlator -> chain::correlate;  plot::waterfall core -> chain::spectrogram
//! Errors: the reference panics (assert_eq!, unwrap, expect); every non-zero ae_status becomes a panic carrying the
//! library's message, which is the reference's text where the reference has one ("Vectors must have same length").
//! Threading: handles are neither Send nor Sync — the per-device context behind them is not locked.
use aether_b200_sys as sys;
use aether_primitives::{cf32, fft::{Fft, Scale}, modulation::Modulation, vecops::VecOps};
use std::ffi::{CStr, CString};
use std::os::raw::{c_int, c_void};
use std::ptr::null_mut;

fn ck(st: sys::ae_status) {
    if st != sys::AE_OK {
        let msg = unsafe { CStr::from_ptr(sys::ae_last_error_string()) };
        panic!("{}", msg.to_string_lossy());
    }
}
fn scale_args(s: Scale) -> (c_int, f32) {
    match s { Scale::None => (sys::AE_SCALE_NONE, 1.0), Scale::SN => (sys::AE_SCALE_SN, 1.0), Scale::N => (sys::AE_SCALE_N, 1.0), Scale::X(x) => (sys::AE_SCALE_X, x) }
}
/// `Scale::scale` factor as the reference computes it (src/fft.rs:22-37), evaluated by the library
pub fn scale_factor(s: Scale, n: usize) -> f32 { let (k, x) = scale_args(s); let mut f = 0f32; unsafe { ck(sys::ae_scale_factor(k, n, x, &mut f)) }; f }

/// reference quirks F3/F4/F5: `Reference` reproduces what the code does, `Corrected` what its docs say
#[derive(Clone, Copy, PartialEq)] pub enum Compat { Reference, Corrected }
impl Compat { fn raw(self) -> c_int { match self { Compat::Reference => sys::AE_COMPAT_REFERENCE, Compat::Corrected => sys::AE_COMPAT_CORRECTED } } }

// ---------------------------------------------------------------- runtime
pub mod runtime {
    use super::*;
    pub fn version() -> String { unsafe { CStr::from_ptr(sys::ae_version()) }.to_string_lossy().into_owned() }
    pub fn device_count() -> i32 { let mut n = 0; unsafe { ck(sys::ae_device_count(&mut n)) }; n }
    pub fn init(device: i32) { unsafe { ck(sys::ae_init(device)) } }
    pub fn set_stream(cuda_stream: *mut c_void) { unsafe { ck(sys::ae_set_stream(cuda_stream)) } }
    pub fn get_stream() -> *mut c_void { unsafe { sys::ae_get_stream() } }
    pub fn sync() { unsafe { ck(sys::ae_sync()) } }
    pub fn sm_count() -> i32 { let mut n = 0; unsafe { ck(sys::ae_sm_count(&mut n)) }; n }
    pub fn launch_count() -> u64 { unsafe { sys::ae_launch_count() } }
    /// Philox4x32-10 block (the AWGN generator's counter-based RNG), host side: known-answer checks
    pub fn philox4x32_10(ctr: [u32; 4], key: [u32; 2]) -> [u32; 4] { let mut o = [0u32; 4]; unsafe { ck(sys::ae_philox4x32_10(ctr.as_ptr(), key.as_ptr(), o.as_mut_ptr())) }; o }
    /// CUDA graph over the context stream: `Graph::record(|| { ...library calls... })` records the launches instead of
    /// running them, `launch()` replays them with one driver call (launch-bound shapes: the 1M-symbol modem loop-back)
    pub struct Graph { h: *mut sys::ae_graph }
    impl Graph {
        pub fn record(f: impl FnOnce()) -> Self { unsafe { ck(sys::ae_graph_begin()) }; f(); let mut h = null_mut(); unsafe { ck(sys::ae_graph_end(&mut h)) }; Graph { h } }
        pub fn launch(&mut self) { unsafe { ck(sys::ae_graph_launch(self.h)) } }
    }
    impl Drop for Graph { fn drop(&mut self) { unsafe { sys::ae_graph_destroy(self.h); } } }
    /// page-locked host buffer for the host pipelines (`Pipe`, `FftFirDemod::run_host`)
    pub struct PinnedBuf { pub ptr: *mut c_void, pub bytes: usize }
    impl PinnedBuf { pub fn new(bytes: usize) -> Self { let mut p = null_mut(); unsafe { ck(sys::ae_host_alloc(bytes, &mut p)) }; PinnedBuf { ptr: p, bytes } } }
    impl Drop for PinnedBuf { fn drop(&mut self) { unsafe { sys::ae_host_free(self.ptr); } } }
}

// ---------------------------------------------------------------- buffers
/// Device-resident `Vec<cf32>`: len and capacity behave like Vec's (append semantics of interpolate / fill / demod)
pub struct DeviceVec { h: *mut sys::ae_vec }
impl DeviceVec {
    pub fn with_capacity(cap: usize) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_vec_alloc(0, cap, &mut h)) }; DeviceVec { h } }
    pub fn zeros(len: usize) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_vec_alloc(len, len, &mut h)); ck(sys::ae_vec_zero(h)) }; DeviceVec { h } }
    pub fn from_slice(v: &[cf32]) -> Self {
        let mut h = null_mut();
        unsafe { ck(sys::ae_vec_alloc(v.len(), v.len(), &mut h)); ck(sys::ae_vec_upload(h, v.as_ptr() as *const _, v.len())); }
        DeviceVec { h }
    }
    /// borrow foreign device memory (`len` cf32 at `device_ptr`); pending ops are flushed when the handle drops
    pub unsafe fn wrap(device_ptr: *mut c_void, len: usize) -> Self { let mut h = null_mut(); ck(sys::ae_vec_wrap(device_ptr, len, &mut h)); DeviceVec { h } }
    /// `&mut v[offset..offset + len]`
    pub fn view(&mut self, offset: usize, len: usize) -> DeviceVec { let mut h = null_mut(); unsafe { ck(sys::ae_vec_view(self.h, offset, len, &mut h)) }; DeviceVec { h } }
    pub fn len(&self) -> usize { unsafe { sys::ae_vec_len(self.h) } }
    pub fn capacity(&self) -> usize { unsafe { sys::ae_vec_capacity(self.h) } }
    pub fn set_len(&mut self, len: usize) { unsafe { ck(sys::ae_vec_set_len(self.h, len)) } }
    pub fn clear(&mut self) { self.set_len(0) }
    pub fn reserve(&mut self, cap: usize) { unsafe { ck(sys::ae_vec_reserve(self.h, cap)) } }
    pub fn device_ptr(&mut self) -> *mut c_void { let mut p = null_mut(); unsafe { ck(sys::ae_vec_device_ptr(self.h, &mut p)) }; p }
    pub fn upload(&mut self, v: &[cf32]) { unsafe { ck(sys::ae_vec_upload(self.h, v.as_ptr() as *const _, v.len())) } }
    pub fn download(&mut self, out: &mut [cf32]) { unsafe { ck(sys::ae_vec_download(self.h, out.as_mut_ptr() as *mut _, out.len())) } }
    pub fn to_vec(&mut self) -> Vec<cf32> { let mut out = vec![cf32::default(); self.len()]; self.download(&mut out); out }
    /// run the recorded VecOps tape now (one fused kernel); downloads and consumers do it implicitly
    pub fn flush(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_flush(self.h)) }; self }
    pub fn pending_ops(&self) -> usize { unsafe { sys::ae_vec_pending_ops(self.h) } }
    /// `Scale::scale(&mut data)` (src/fft.rs:22-37)
    pub fn scale_kind(&mut self, s: Scale) -> &mut Self { let (k, x) = scale_args(s); unsafe { ck(sys::ae_vec_scale_kind(self.h, k, x)) }; self }
    /// README TODO "VecStats": min/max (index), mean, power
    pub fn vec_stats(&mut self) -> sys::ae_vecstats { let mut s = std::mem::MaybeUninit::<sys::ae_vecstats>::uninit(); unsafe { ck(sys::ae_vec_stats(self.h, s.as_mut_ptr())); s.assume_init() } }
    /// util::file raw format (src/util/file.rs:29-107): native-endian cf32 structs, no header
    pub fn read_raw(path: &str) -> Self { let c = CString::new(path).unwrap(); let mut h = null_mut(); unsafe { ck(sys::ae_vec_read_raw(c.as_ptr(), &mut h)) }; DeviceVec { h } }
    pub fn write_raw(&mut self, path: &str) { let c = CString::new(path).unwrap(); unsafe { ck(sys::ae_vec_write_raw(self.h, c.as_ptr())) } }
    // VecOps with DEVICE operands: same names + `_dev`; chained calls fuse into one kernel (op tape)
    pub fn vec_mul_dev(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_mul(self.h, o.h)) }; self }
    pub fn vec_div_dev(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_div(self.h, o.h)) }; self }
    pub fn vec_add_dev(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_add(self.h, o.h)) }; self }
    pub fn vec_sub_dev(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_sub(self.h, o.h)) }; self }
    pub fn vec_clone_dev(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_clone(self.h, o.h)) }; self }
    pub fn vec_rfft_dev(&mut self, fft: &mut CudaFft, s: Scale) -> &mut Self { fft.ifwd_dev(self, s, 1); self }
    pub fn vec_rifft_dev(&mut self, fft: &mut CudaFft, s: Scale) -> &mut Self { fft.ibwd_dev(self, s, 1); self }
}
impl Drop for DeviceVec { fn drop(&mut self) { unsafe { sys::ae_vec_free(self.h); } } }

/// `impl VecOps` with the trait's own signatures: `other: impl AsRef<[cf32]>` operands are HOST slices and are
/// uploaded for the call; use the `_dev` methods above to keep operands on the device.
impl VecOps for DeviceVec {
    fn vec_scale(&mut self, s: f32) -> &mut Self { unsafe { ck(sys::ae_vec_scale(self.h, s)) }; self }                   // :94-97
    fn vec_mul(&mut self, o: impl AsRef<[cf32]>) -> &mut Self { let t = DeviceVec::from_slice(o.as_ref()); self.vec_mul_dev(&t).flush() }   // :99-112, AE_ELEN -> "Vectors must have same length"
    fn vec_div(&mut self, o: impl AsRef<[cf32]>) -> &mut Self { let t = DeviceVec::from_slice(o.as_ref()); self.vec_div_dev(&t).flush() }   // :114-125
    fn vec_conj(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_conj(self.h)) }; self }                                  // :127-130
    fn vec_mirror(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_mirror(self.h)) }; self }                              // :157-161
    fn vec_clone(&mut self, o: impl AsRef<[cf32]>) -> &mut Self { let t = DeviceVec::from_slice(o.as_ref()); self.vec_clone_dev(&t).flush() } // :163-172
    fn vec_zero(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_zero(self.h)) }; self }                                  // :174-177
    fn vec_mutate(&mut self, mut f: impl FnMut(&mut cf32)) -> &mut Self {                                                  // :179-182: host round trip (slow path)
        extern "C" fn tramp<F: FnMut(&mut cf32)>(e: *mut sys::ae_cf32, u: *mut c_void) { unsafe { (*(u as *mut F))(&mut *(e as *mut cf32)) } }
        fn call<F: FnMut(&mut cf32)>(h: *mut sys::ae_vec, f: &mut F) { unsafe { ck(sys::ae_vec_mutate(h, tramp::<F>, f as *mut F as *mut c_void)) } }
        call(self.h, &mut f);
        self
    }
    fn vec_add(&mut self, o: impl AsRef<[cf32]>) -> &mut Self { let t = DeviceVec::from_slice(o.as_ref()); self.vec_add_dev(&t).flush() }   // :132-142
    fn vec_sub(&mut self, o: impl AsRef<[cf32]>) -> &mut Self { let t = DeviceVec::from_slice(o.as_ref()); self.vec_sub_dev(&t).flush() }   // :144-155
    fn vec_fft(&mut self, s: Scale) -> &mut Self { let (k, x) = scale_args(s); unsafe { ck(sys::ae_vec_fft(self.h, k, x, sys::AE_COMPAT_REFERENCE)) }; self }   // :301-306
    fn vec_ifft(&mut self, s: Scale) -> &mut Self { let (k, x) = scale_args(s); unsafe { ck(sys::ae_vec_ifft(self.h, k, x, sys::AE_COMPAT_REFERENCE)) }; self } // :308-313
    /// any `impl Fft`: the trait only offers host slices, so this is download -> fft.ifwd -> upload; `vec_rfft_dev` stays on the device
    fn vec_rfft(&mut self, fft: &mut impl Fft, s: Scale) -> &mut Self { let mut h = self.to_vec(); fft.ifwd(&mut h, s); self.upload(&h); self }   // :315-319
    fn vec_rifft(&mut self, fft: &mut impl Fft, s: Scale) -> &mut Self { let mut h = self.to_vec(); fft.ibwd(&mut h, s); self.upload(&h); self }  // :321-325
}

/// Device `Vec<u8>`, one byte per bit (src/modulation.rs:102-103)
pub struct DeviceBits { h: *mut sys::ae_bits }
impl DeviceBits {
    pub fn with_capacity(cap: usize) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_bits_alloc(0, cap, &mut h)) }; DeviceBits { h } }
    pub fn from_slice(b: &[u8]) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_bits_alloc(b.len(), b.len(), &mut h)); ck(sys::ae_bits_upload(h, b.as_ptr(), b.len())) }; DeviceBits { h } }
    pub unsafe fn wrap(device_ptr: *mut c_void, len: usize) -> Self { let mut h = null_mut(); ck(sys::ae_bits_wrap(device_ptr, len, &mut h)); DeviceBits { h } }
    pub fn len(&self) -> usize { unsafe { sys::ae_bits_len(self.h) } }
    pub fn capacity(&self) -> usize { unsafe { sys::ae_bits_capacity(self.h) } }
    pub fn set_len(&mut self, len: usize) { unsafe { ck(sys::ae_bits_set_len(self.h, len)) } }
    pub fn device_ptr(&mut self) -> *mut c_void { let mut p = null_mut(); unsafe { ck(sys::ae_bits_device_ptr(self.h, &mut p)) }; p }
    pub fn to_vec(&mut self) -> Vec<u8> { let mut o = vec![0u8; self.len()]; unsafe { ck(sys::ae_bits_download(self.h, o.as_mut_ptr(), o.len())) }; o }
}
impl Drop for DeviceBits { fn drop(&mut self) { unsafe { sys::ae_bits_free(self.h); } } }

/// Device `Vec<f32>` (spectrogram levels)
pub struct DeviceF32 { h: *mut sys::ae_f32 }
impl DeviceF32 {
    pub fn new(len: usize) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_f32_alloc(len, &mut h)) }; DeviceF32 { h } }
    pub fn len(&self) -> usize { unsafe { sys::ae_f32_len(self.h) } }
    pub fn device_ptr(&mut self) -> *mut c_void { let mut p = null_mut(); unsafe { ck(sys::ae_f32_device_ptr(self.h, &mut p)) }; p }
    pub fn to_vec(&mut self) -> Vec<f32> { let mut o = vec![0f32; self.len()]; unsafe { ck(sys::ae_f32_download(self.h, o.as_mut_ptr(), o.len())) }; o }
    pub fn vec_stats(&mut self) -> sys::ae_vecstats { let mut s = std::mem::MaybeUninit::<sys::ae_vecstats>::uninit(); unsafe { ck(sys::ae_f32_stats(self.h, s.as_mut_ptr())); s.assume_init() } }
}
impl Drop for DeviceF32 { fn drop(&mut self) { unsafe { sys::ae_f32_free(self.h); } } }

// ---------------------------------------------------------------- fft
/// `Cfft` (src/fft.rs:134-235).  Under feature `fft_b200` the crate re-exports it as `fft::Cfft`.
pub struct CudaFft { h: *mut sys::ae_fft, len: usize, io: DeviceVec, tmp: Vec<cf32> }
impl CudaFft {
    pub fn with_len(len: usize) -> Self {                                                                                  // :147-159
        let mut h = null_mut();
        unsafe { ck(sys::ae_fft_create(len, &mut h)) };
        debug_assert_eq!(unsafe { sys::ae_fft_len(h) }, len);
        CudaFft { h, len, io: DeviceVec::zeros(len), tmp: vec![cf32::default(); len] }
    }
    /// `Reference`: Cfft::fwd uses the e^{+} exponent like the crate (F3: FFTplanner::new(true)); `Corrected`: e^{-}
    pub fn set_compat(&mut self, c: Compat) { unsafe { ck(sys::ae_fft_set_compat(self.h, c.raw())) } }
    /// batched in-place transforms on the device: `howmany` frames of `len` samples
    pub fn ifwd_dev(&mut self, v: &mut DeviceVec, s: Scale, howmany: usize) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, sys::AE_FFT_FWD, v.h, null_mut(), k, x, howmany)) } }
    pub fn ibwd_dev(&mut self, v: &mut DeviceVec, s: Scale, howmany: usize) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, sys::AE_FFT_BWD, v.h, null_mut(), k, x, howmany)) } }
    pub fn fwd_dev(&mut self, i: &mut DeviceVec, o: &mut DeviceVec, s: Scale, howmany: usize) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, sys::AE_FFT_FWD, i.h, o.h, k, x, howmany)) } }
    pub fn bwd_dev(&mut self, i: &mut DeviceVec, o: &mut DeviceVec, s: Scale, howmany: usize) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, sys::AE_FFT_BWD, i.h, o.h, k, x, howmany)) } }
    fn host(&mut self, dir: c_int, input: &[cf32], output: &mut [cf32], s: Scale) {
        assert_eq!(self.len, input.len(), "Input and FFT must be the same length");                                        // :163-167
        assert_eq!(self.len, output.len(), "Input and FFT must be the same length");
        let (k, x) = scale_args(s);
        unsafe {
            ck(sys::ae_vec_upload(self.io.h, input.as_ptr() as *const _, input.len()));
            ck(sys::ae_fft_exec(self.h, dir, self.io.h, null_mut(), k, x, 1));
            ck(sys::ae_vec_download(self.io.h, output.as_mut_ptr() as *mut _, output.len()));
        }
    }
    /// tfwd/tbwd: the result lives in the plan's scratch (valid until the next call on this plan), copied into `self.tmp`
    fn host_tmp(&mut self, dir: c_int, input: &[cf32], s: Scale) -> &[cf32] {
        assert_eq!(self.len, input.len(), "Input and FFT must be the same length");
        let (k, x) = scale_args(s);
        let mut view = null_mut();
        unsafe {
            ck(sys::ae_vec_upload(self.io.h, input.as_ptr() as *const _, input.len()));
            ck(sys::ae_fft_exec_tmp(self.h, dir, self.io.h, k, x, 1, &mut view));
            ck(sys::ae_vec_download(view, self.tmp.as_mut_ptr() as *mut _, self.len));                                      // plan-owned view: not freed
        }
        &self.tmp
    }
}
/// The crate's own `Fft` trait on HOST slices: upload, transform, download — `impl Fft for Cfft` (src/fft.rs:161-235)
impl Fft for CudaFft {
    fn fwd(&mut self, input: &[cf32], output: &mut [cf32], s: Scale) { self.host(sys::AE_FFT_FWD, input, output, s) }       // :162-171
    fn bwd(&mut self, input: &[cf32], output: &mut [cf32], s: Scale) { self.host(sys::AE_FFT_BWD, input, output, s) }       // :173-182
    fn ifwd(&mut self, input: &mut [cf32], s: Scale) { let t = input.to_vec(); self.host(sys::AE_FFT_FWD, &t, input, s) }   // :184-193
    fn ibwd(&mut self, input: &mut [cf32], s: Scale) { let t = input.to_vec(); self.host(sys::AE_FFT_BWD, &t, input, s) }   // :195-204
    fn tfwd(&mut self, input: &[cf32], s: Scale) -> &[cf32] { self.host_tmp(sys::AE_FFT_FWD, input, s) }                    // :206-217
    fn tbwd(&mut self, input: &[cf32], s: Scale) -> &[cf32] { self.host_tmp(sys::AE_FFT_BWD, input, s) }                    // :219-230
    fn len(&self) -> usize { self.len }
}
impl Drop for CudaFft { fn drop(&mut self) { unsafe { sys::ae_fft_destroy(self.h); } } }

// ---------------------------------------------------------------- fir
#[derive(Clone, Copy)] pub enum FirMode { Auto, Direct, OverlapSave }
/// src/fir.rs:3-22 holds a constructor only; `filter` is y[n] = sum_k taps[k] x[n-k], output length = input length
pub struct Fir { h: *mut sys::ae_fir }
impl Fir {
    pub fn new(taps: &[cf32], _input_len: usize) -> Self { Self::with_mode(taps, FirMode::Auto) }                          // Fir::new (src/fir.rs:14)
    pub fn with_mode(taps: &[cf32], m: FirMode) -> Self {
        let mode = match m { FirMode::Auto => sys::AE_FIR_AUTO, FirMode::Direct => sys::AE_FIR_DIRECT, FirMode::OverlapSave => sys::AE_FIR_OVERLAP_SAVE };
        let mut h = null_mut();
        unsafe { ck(sys::ae_fir_create(taps.as_ptr() as *const _, taps.len(), mode, &mut h)) };
        Fir { h }
    }
    pub fn ntaps(&self) -> usize { unsafe { sys::ae_fir_ntaps(self.h) } }
    /// outputs per overlap-save segment (1 for the direct form): shard starts on multiples of it are bit-identical to the unsharded stream
    pub fn block_hop(&self) -> usize { unsafe { sys::ae_fir_block_hop(self.h) } }
    pub fn reset(&mut self) { unsafe { ck(sys::ae_fir_reset(self.h)) } }
    /// frame_len = 0: streaming (ntaps-1 samples of history carried across calls); > 0: zero state at every frame start
    pub fn filter_dev(&mut self, input: &mut DeviceVec, output: &mut DeviceVec, frame_len: usize) { unsafe { ck(sys::ae_fir_exec(self.h, input.h, output.h, frame_len)) } }
    pub fn filter(&mut self, input: &[cf32], output: &mut [cf32]) { let mut i = DeviceVec::from_slice(input); let mut o = DeviceVec::zeros(input.len()); self.filter_dev(&mut i, &mut o, 0); o.download(output) }
}
impl Drop for Fir { fn drop(&mut self) { unsafe { sys::ae_fir_destroy(self.h); } } }

// ---------------------------------------------------------------- sampling
pub mod sampling {
    use super::*;
    /// src/sampling.rs:7-24 (appends to dst; `Reference` keeps the `im: x1.re + ...` quirk F4)
    pub fn interpolate_dev(src: &mut DeviceVec, dst: &mut DeviceVec, n_between: usize, c: Compat) { unsafe { ck(sys::ae_interpolate(src.h, dst.h, n_between, c.raw())) } }
    pub fn interpolate(src: &[cf32], dst: &mut Vec<cf32>, n_between: usize) {
        let (mut s, mut d) = (DeviceVec::from_slice(src), DeviceVec::with_capacity((src.len().max(1) - 1) * (n_between + 1) + 1));
        interpolate_dev(&mut s, &mut d, n_between, Compat::Reference);
        dst.extend(d.to_vec());
    }
    /// src/sampling.rs:28-42; `strict` turns the reference's debug_assert (divisibility) into an error
    pub fn downsample_dev(src: &mut DeviceVec, dst: &mut DeviceVec, strict: bool) { unsafe { ck(sys::ae_downsample(src.h, dst.h, strict as c_int)) } }
    /// src/sampling.rs:49-62 (same result, the reference's step_by variant)
    pub fn downsample_sb_dev(src: &mut DeviceVec, dst: &mut DeviceVec, strict: bool) { unsafe { ck(sys::ae_downsample_sb(src.h, dst.h, strict as c_int)) } }
    /// the generic `T: Copy` of the reference for the one other element type of the path: bits
    pub fn downsample_bits_dev(src: &mut DeviceBits, dst: &mut DeviceBits, strict: bool) { unsafe { ck(sys::ae_downsample_bits(src.h, dst.h, strict as c_int)) } }
    pub fn downsample(src: &[cf32], dst: &mut [cf32]) { let (mut s, mut d) = (DeviceVec::from_slice(src), DeviceVec::zeros(dst.len())); downsample_dev(&mut s, &mut d, false); d.download(dst) }
    pub fn downsample_sb(src: &[cf32], dst: &mut [cf32]) { let (mut s, mut d) = (DeviceVec::from_slice(src), DeviceVec::zeros(dst.len())); downsample_sb_dev(&mut s, &mut d, false); d.download(dst) }
}

// ---------------------------------------------------------------- modulation
/// Device constellation table; `Bpsk` / `Qpsk` below implement the crate's `Modulation` trait with it
pub struct DeviceMod { h: *mut sys::ae_mod }
impl DeviceMod {
    pub fn bpsk() -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_mod_bpsk(&mut h)) }; DeviceMod { h } }            // table src/modulation.rs:77
    pub fn qpsk() -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_mod_qpsk(&mut h)) }; DeviceMod { h } }            // table :87-92
    pub fn from_table(t: &[cf32]) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_mod_create(t.as_ptr() as *const _, t.len(), &mut h)) }; DeviceMod { h } }
    pub fn bits_per_symbol(&self) -> usize { unsafe { sys::ae_mod_bits_per_symbol(self.h) } }
    pub fn modulate_dev(&self, bits: &mut DeviceBits, out: &mut DeviceVec) { unsafe { ck(sys::ae_mod_modulate(self.h, bits.h, out.h)) } }        // :115-121 (appends)
    pub fn modulate_into_dev(&self, bits: &mut DeviceBits, out: &mut DeviceVec) { unsafe { ck(sys::ae_mod_modulate_into(self.h, bits.h, out.h)) } } // :123-131 (truncates like zip)
    pub fn demod_naive_dev(&self, sym: &mut DeviceVec, out: &mut DeviceBits, c: Compat) { unsafe { ck(sys::ae_mod_demod(self.h, sym.h, out.h, c.raw())) } } // :133-144, :33-56 (appends)
}
impl Drop for DeviceMod { fn drop(&mut self) { unsafe { sys::ae_mod_destroy(self.h); } } }

macro_rules! impl_modulation {
    ($name:ident, $bps:expr, $ctor:ident, $table:expr) => {
        pub struct $name { dev: DeviceMod }
        impl $name { pub fn new() -> Self { $name { dev: DeviceMod::$ctor() } } pub fn device(&self) -> &DeviceMod { &self.dev } }
        impl Modulation for $name {
            const BITS_PER_SYMBOL: usize = $bps;
            fn symbol(&self, idx: usize) -> cf32 { $table[idx] }
            fn modulate(&self, input: &[u8]) -> Vec<cf32> {
                let (mut b, mut o) = (DeviceBits::from_slice(input), DeviceVec::with_capacity(input.len() / $bps + 1));
                self.dev.modulate_dev(&mut b, &mut o);                       // AE_EIDX on a ragged tail = the reference's index-out-of-bounds panic
                o.to_vec()
            }
            fn modulate_into<'a>(&self, input: &[u8], output: &mut impl Iterator<Item = &'a mut cf32>) {
                for (s, out) in self.modulate(&input[..input.len() / $bps * $bps]).into_iter().zip(output) { *out = s }
            }
            fn demod_naive<'a>(&self, symbols: &mut impl Iterator<Item = &'a cf32>, output: &mut Vec<u8>) {
                let host: Vec<cf32> = symbols.cloned().collect();
                let (mut s, mut b) = (DeviceVec::from_slice(&host), DeviceBits::with_capacity(host.len() * $bps));
                self.dev.demod_naive_dev(&mut s, &mut b, Compat::Reference);  // QPSK pushes idx & 2 in {0,2} like the reference (F5a)
                output.extend(b.to_vec());
            }
        }
    };
}
impl_modulation!(Bpsk, 1, bpsk, [cf32 { re: 1.0, im: 1.0 }, cf32 { re: -1.0, im: -1.0 }]);
impl_modulation!(Qpsk, 2, qpsk, [cf32 { re: 1.0, im: 1.0 }, cf32 { re: -1.0, im: 1.0 }, cf32 { re: 1.0, im: -1.0 }, cf32 { re: -1.0, im: -1.0 }]);

// ---------------------------------------------------------------- noise
pub mod noise {
    use super::*;
    /// src/noise.rs:20-71 on Philox4x32-10 + Box-Muller instead of ChaCha20 + ziggurat (validated statistically)
    pub struct Awgn { pub power: f32, h: *mut sys::ae_awgn }
    /// default seed 815, power 1 (src/noise.rs:9-11)
    pub fn generator() -> Awgn { let mut h = null_mut(); unsafe { ck(sys::ae_awgn_generator(&mut h)) }; Awgn { power: 1.0, h } }
    pub fn new(power: f32, seed: u64) -> Awgn { Awgn::new(power, seed) }                                                   // :14-16
    impl Awgn {
        pub fn new(power: f32, seed: u64) -> Awgn { let mut h = null_mut(); unsafe { ck(sys::ae_awgn_create(power, seed, &mut h)) }; Awgn { power, h } }   // :29-37
        pub fn set_power(&mut self, power: f32) { self.power = power; unsafe { ck(sys::ae_awgn_set_power(self.h, power)) } }   // :47-50
        /// independent stream per frame / channel / GPU: results do not depend on how the work is split
        pub fn set_stream_id(&mut self, id: u64) { unsafe { ck(sys::ae_awgn_set_stream_id(self.h, id)) } }
        pub fn seek(&mut self, sample_offset: u64) { unsafe { ck(sys::ae_awgn_seek(self.h, sample_offset)) } }
        pub fn tell(&self) -> u64 { unsafe { sys::ae_awgn_tell(self.h) } }
        pub fn apply_dev(&mut self, signal: &mut DeviceVec, c: Compat) { unsafe { ck(sys::ae_awgn_apply(self.h, signal.h, c.raw())) } }
        pub fn fill_dev(&mut self, target: &mut DeviceVec) { unsafe { ck(sys::ae_awgn_fill(self.h, target.h)) } }
        pub fn apply(&mut self, signal: &mut [cf32]) { let mut d = DeviceVec::from_slice(signal); self.apply_dev(&mut d, Compat::Reference); d.download(signal) }   // :53-59 (sigma = power, F5b)
        pub fn fill(&mut self, target: &mut Vec<cf32>) {                                                                      // :62-66: push until len == capacity
            let n = target.capacity() - target.len();
            let mut d = DeviceVec::with_capacity(n);
            self.fill_dev(&mut d);
            target.extend(d.to_vec());
        }
        pub fn iter(&mut self) -> NoiseIter { NoiseIter { noisegen: self, buf: Vec::new(), pos: 0 } }                         // :68-70
        pub(crate) fn raw(&self) -> *mut sys::ae_awgn { self.h }
    }
    impl Drop for Awgn { fn drop(&mut self) { unsafe { sys::ae_awgn_destroy(self.h); } } }
    /// infinite iterator (src/noise.rs:73-84), refilled 4096 samples at a time from the device stream
    pub struct NoiseIter<'a> { noisegen: &'a mut Awgn, buf: Vec<cf32>, pos: usize }
    impl<'a> Iterator for NoiseIter<'a> {
        type Item = cf32;
        fn next(&mut self) -> Option<cf32> {
            if self.pos == self.buf.len() {
                self.buf.resize(4096, cf32::default());
                unsafe { ck(sys::ae_awgn_next_host(self.noisegen.h, self.buf.as_mut_ptr() as *mut _, self.buf.len())) };
                self.pos = 0;
            }
            self.pos += 1;
            Some(self.buf[self.pos - 1])
        }
    }
}

// ---------------------------------------------------------------- sequence
pub mod sequence {
    use super::*;
    pub fn expand_dev(seed: u64, len: usize, out: &mut DeviceBits) { unsafe { ck(sys::ae_mseq_expand(seed, len, out.h)) } }
    pub fn expand(seed: u64, len: usize) -> Vec<u8> { let mut o = DeviceBits::with_capacity(len); expand_dev(seed, len, &mut o); o.to_vec() }   // src/sequence.rs:18-21
    /// `generate` (src/sequence.rs:47-53) for generators of the crate's documented form x[n] = (sum_t x[n - back[t]]) % 2 (:42);
    /// an arbitrary closure cannot cross the device boundary
    pub fn generate_taps_dev(init: &[u8], back_offsets: &[u32], len: usize, out: &mut DeviceBits) {
        unsafe { ck(sys::ae_mseq_generate(init.as_ptr(), init.len(), back_offsets.as_ptr(), back_offsets.len(), len, out.h)) }
    }
    pub fn generate_taps(init: Vec<u8>, back_offsets: &[u32], len: usize) -> Vec<u8> { let mut o = DeviceBits::with_capacity(len); generate_taps_dev(&init, back_offsets, len, &mut o); o.to_vec() }
}

// ---------------------------------------------------------------- statistics + the one collective
/// {bit_errors, n_bits, sum|e|^2, sum|r|^2} on the device: the only quantities ever reduced across GPUs
pub struct DeviceStats { h: *mut sys::ae_stats }
impl DeviceStats {
    pub fn new() -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_stats_alloc(&mut h)) }; DeviceStats { h } }
    pub fn zero(&mut self) { unsafe { ck(sys::ae_stats_zero(self.h)) } }
    pub fn read(&self) -> sys::ae_stats { let mut s = sys::ae_stats { bit_errors: 0, n_bits: 0, err_pow: 0.0, ref_pow: 0.0 }; unsafe { ck(sys::ae_stats_read(self.h, &mut s)) }; s }
    pub fn count_bit_errors(&mut self, a: &mut DeviceBits, b: &mut DeviceBits) { unsafe { ck(sys::ae_count_bit_errors(a.h, b.h, self.h)) } }
    pub fn evm_accumulate(&mut self, actual: &mut DeviceVec, reference: &mut DeviceVec) { unsafe { ck(sys::ae_evm_accumulate(actual.h, reference.h, self.h)) } }
    /// EVM in dB as the macro's doc defines it: 10 log10(P_error / P_ref) (src/lib.rs:21)
    pub fn evm_db(s: &sys::ae_stats) -> f64 { 10.0 * (s.err_pow / s.ref_pow).log10() }
    /// in place: sum over the communicator's ranks (ncclAllReduce inside the library; asynchronous, `read` synchronises)
    pub fn allreduce(&mut self, comm: &mut Comm) { unsafe { ck(sys::ae_stats_allreduce(self.h, comm.h)) } }
}
impl Drop for DeviceStats { fn drop(&mut self) { unsafe { sys::ae_stats_free(self.h); } } }

/// NCCL communicator owned by the library.  One process per GPU: rank 0 makes the id, hands the 128 bytes to the other
/// ranks (MPI, a file, a socket), every rank joins.  One process, n GPUs: `init_all` + `allreduce_all`.
pub struct Comm { h: *mut sys::ae_comm }
impl Comm {
    pub fn unique_id() -> [u8; sys::AE_COMM_ID_BYTES] { let mut id = [0u8; sys::AE_COMM_ID_BYTES]; unsafe { ck(sys::ae_comm_unique_id(id.as_mut_ptr())) }; id }
    pub fn init_rank(id: &[u8; sys::AE_COMM_ID_BYTES], nranks: i32, rank: i32) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_comm_init_rank(id.as_ptr(), nranks, rank, &mut h)) }; Comm { h } }
    pub fn init_all(ndev: i32) -> Vec<Comm> { let mut hs = vec![null_mut(); ndev as usize]; unsafe { ck(sys::ae_comm_init_all(ndev, hs.as_mut_ptr())) }; hs.into_iter().map(|h| Comm { h }).collect() }
    /// (nranks, rank, device)
    pub fn info(&self) -> (i32, i32, i32) { let (mut n, mut r, mut d) = (0, 0, 0); unsafe { ck(sys::ae_comm_info(self.h, &mut n, &mut r, &mut d)) }; (n, r, d) }
    /// every device's reduction inside one NCCL group (single-process form)
    pub fn allreduce_all(stats: &mut [DeviceStats], comms: &mut [Comm]) {
        assert_eq!(stats.len(), comms.len(), "Vectors must have same length");
        let mut s: Vec<_> = stats.iter().map(|x| x.h).collect();
        let mut c: Vec<_> = comms.iter().map(|x| x.h).collect();
        unsafe { ck(sys::ae_stats_allreduce_all(s.as_mut_ptr(), c.as_mut_ptr(), s.len() as c_int)) }
    }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { sys::ae_comm_destroy(self.h); } } }

// ---------------------------------------------------------------- fused chains
pub mod chain {
    use super::*;
    /// examples/modem.rs:15-32 in one kernel: modulate -> Awgn::apply -> demod_naive (+ bit-error count into `stats`)
    pub fn modem_fused(m: &DeviceMod, g: &mut noise::Awgn, bits_in: &mut DeviceBits, bits_out: &mut DeviceBits, stats: Option<&mut DeviceStats>, c: Compat) {
        unsafe { ck(sys::ae_modem_fused(m.h, g.raw(), bits_in.h, bits_out.h, stats.map_or(null_mut(), |s| s.h), c.raw())) }
    }
    /// BASELINE config 5: M-sequence -> QPSK -> bwd FFT(SN) -> AWGN -> fwd FFT(SN) -> demod -> BER/EVM, frames
    /// [first_frame_id, first_frame_id + frames) of the global job (results do not depend on the split over GPUs)
    pub fn ofdm_chain(fft_len: usize, frames: usize, first_frame_id: u64, noise_power: f32, noise_seed: u64, c: Compat,
                      tx_bits: Option<&mut DeviceBits>, rx_bits: Option<&mut DeviceBits>, stats: &mut DeviceStats) {
        unsafe { ck(sys::ae_ofdm_chain(fft_len, frames, first_frame_id, noise_power, noise_seed, c.raw(), tx_bits.map_or(null_mut(), |b| b.h), rx_bits.map_or(null_mut(), |b| b.h), stats.h)) }
    }
    /// util::plot::waterfall's compute core (src/util/plot.rs:46-68): chunks -> rfft(SN) -> mirror -> norm -> dB
    pub fn spectrogram(fft: &mut CudaFft, symbols: &mut DeviceVec, levels: &mut DeviceF32, use_db: bool) { unsafe { ck(sys::ae_spectrogram(fft.h, symbols.h, levels.h, use_db as c_int)) } }
    /// benches/benches.rs:410-416: input.vec_rfft(fft, s).vec_mul(sig).vec_rifft(fft, s) per frame, one kernel
    pub fn correlate(fft: &mut CudaFft, inout: &mut DeviceVec, sig: &mut DeviceVec, s: Scale, howmany: usize) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_correlate(fft.h, inout.h, sig.h, k, x, howmany)) } }

    /// The headline chain: per frame Cfft::fwd(scale) -> FIR (zero state per frame) -> QPSK demod_naive, one kernel
    pub struct FftFirDemod { h: *mut sys::ae_chain, n: usize }
    impl FftFirDemod {
        pub fn new(fft_len: usize, taps: &[cf32], s: Scale, c: Compat) -> Self {
            let (k, x) = scale_args(s);
            let mut h = null_mut();
            unsafe { ck(sys::ae_chain_create(fft_len, taps.as_ptr() as *const _, taps.len(), k, x, c.raw(), &mut h)) };
            FftFirDemod { h, n: fft_len }
        }
        pub fn run(&mut self, input: &mut DeviceVec, bits_out: &mut DeviceBits) { unsafe { ck(sys::ae_chain_exec(self.h, input.h, bits_out.h)) } }
        /// composition of the stand-alone kernels (any frame length); optionally keeps the filtered symbols
        pub fn run_unfused(&mut self, input: &mut DeviceVec, bits_out: &mut DeviceBits, symbols: Option<&mut DeviceVec>) { unsafe { ck(sys::ae_chain_exec_unfused(self.h, input.h, bits_out.h, symbols.map_or(null_mut(), |s| s.h))) } }
        /// host buffers in, host bits out: pinned chunks, H2D / kernel / D2H overlapped on three streams
        pub fn run_host(&mut self, input: &[cf32], bits_out: &mut [u8]) {
            assert_eq!(bits_out.len(), 2 * input.len(), "Vectors must have same length");
            unsafe { ck(sys::ae_chain_exec_host(self.h, input.as_ptr() as *const _, input.len(), bits_out.as_mut_ptr())) }
        }
        pub fn fft_len(&self) -> usize { self.n }
    }
    impl Drop for FftFirDemod { fn drop(&mut self) { unsafe { sys::ae_chain_destroy(self.h); } } }

    /// pipeline.rs + pool.rs for this path (src/pipeline.rs:26-137, src/pool.rs:43-130): host blocks stream through a ring
    /// of `depth` device slots (H2D copy -> fused kernel -> D2H copy), results come back in order, per-stage report
    pub struct Pipe<'a> { h: *mut sys::ae_pipe, _chain: &'a mut FftFirDemod }
    impl<'a> Pipe<'a> {
        pub fn new(chain: &'a mut FftFirDemod, block_frames: usize, depth: i32) -> Self { let mut h = null_mut(); unsafe { ck(sys::ae_pipe_create(chain.h, block_frames, depth, &mut h)) }; Pipe { h, _chain: chain } }
        /// queue one block; `host_in` / `host_bits` must stay valid (and should be pinned) until the block is received
        pub unsafe fn send(&mut self, host_in: *const cf32, host_bits: *mut u8) { ck(sys::ae_pipe_send(self.h, host_in as *const _, host_bits)) }
        /// the bit buffer of the oldest finished block
        pub fn recv(&mut self) -> *mut u8 { let mut p = null_mut(); unsafe { ck(sys::ae_pipe_recv(self.h, &mut p)) }; p }
        pub fn in_flight(&self) -> usize { unsafe { sys::ae_pipe_in_flight(self.h) } }
        /// per stage (h2d, kernel, d2h): processed, active time, rate, utilisation — what pipeline.rs prints (:93-107)
        pub fn report(&mut self, reset: bool) -> [sys::ae_pipe_stage; 3] {
            let mut st = std::mem::MaybeUninit::<[sys::ae_pipe_stage; 3]>::uninit();
            unsafe { ck(sys::ae_pipe_report(self.h, st.as_mut_ptr() as *mut _, reset as c_int)); st.assume_init() }
        }
    }
    impl<'a> Drop for Pipe<'a> { fn drop(&mut self) { unsafe { sys::ae_pipe_destroy(self.h); } } }
}

// ---------------------------------------------------------------- pipeline.rs on CUDA streams
/// `pipeline::new(name, op).add_stage(name, op)....finish()` of the reference (src/pipeline.rs:26-137) with a CUDA stream
/// per stage instead of a thread and CUDA events instead of channels.  `Sender::send` runs every stage's closure once, in
/// order, on the calling thread; the closures only QUEUE work (kernels, `*_async` copies) — while one runs, everything the
/// library launches goes to that stage's stream, so stage k of item i overlaps stage k+1 of item i-1 on the GPU.  Items
/// come out of `Receiver::recv` in the order they went in.  The crate's own `pool::Pool<T>` feeds it unchanged
/// (`T` = a struct of `PinnedBuf`s and `DeviceVec`s).
pub mod pipeline {
    use super::*;
    use std::any::Any;
    use std::cell::RefCell;
    use std::rc::Rc;

    type Boxed = Box<dyn Any>;
    struct Stage { op: Box<dyn FnMut(Boxed) -> Boxed> }
    struct State { h: *mut sys::ae_pipeline, stages: Vec<Box<RefCell<Stage>>> }
    impl Drop for State { fn drop(&mut self) { unsafe { sys::ae_pipeline_destroy(self.h); } } }

    extern "C" fn tramp(user: *mut c_void, _slot: usize, item: *mut c_void) -> sys::ae_status {
        // user: the stage; item: a heap cell holding the travelling value
        let stage = unsafe { &*(user as *const RefCell<Stage>) };
        let cell = unsafe { &mut *(item as *mut Option<Boxed>) };
        // a panic must not unwind through the C frames of ae_pipeline_send: report it as a failed stage
        let r = std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| {
            let v = cell.take().expect("pipeline item");
            *cell = Some((stage.borrow_mut().op)(v));
        }));
        if r.is_ok() { sys::AE_OK } else { sys::AE_EARG }
    }

    pub struct Pipeline<I, O> { st: Rc<RefCell<State>>, _m: std::marker::PhantomData<(I, O)> }
    pub struct Sender<I> { st: Rc<RefCell<State>>, _m: std::marker::PhantomData<I> }
    pub struct Receiver<O> { st: Rc<RefCell<State>>, _m: std::marker::PhantomData<O> }

    fn push_stage<A: 'static, B: 'static>(st: &Rc<RefCell<State>>, name: &str, mut op: impl FnMut(A) -> B + 'static) {
        let stage = Box::new(RefCell::new(Stage { op: Box::new(move |b: Boxed| -> Boxed { Box::new(op(*b.downcast::<A>().expect("stage input type"))) }) }));
        let user = &*stage as *const RefCell<Stage> as *mut c_void;
        let cname = CString::new(name).unwrap();
        let mut s = st.borrow_mut();
        unsafe { ck(sys::ae_pipeline_add_stage(s.h, cname.as_ptr(), tramp, user)) };
        s.stages.push(stage);
    }

    /// `depth` items may be in flight (the reference's channels are unbounded; here the buffers an item carries bound it)
    pub fn new<I: 'static, O: 'static>(name: &str, op: impl FnMut(I) -> O + 'static, depth: i32) -> Pipeline<I, O> {
        let mut h = null_mut();
        unsafe { ck(sys::ae_pipeline_create(depth, &mut h)) };
        let st = Rc::new(RefCell::new(State { h, stages: Vec::new() }));
        push_stage::<I, O>(&st, name, op);
        Pipeline { st, _m: std::marker::PhantomData }
    }
    impl<I: 'static, O: 'static> Pipeline<I, O> {
        pub fn add_stage<U: 'static>(self, name: &str, op: impl FnMut(O) -> U + 'static) -> Pipeline<I, U> {
            push_stage::<O, U>(&self.st, name, op);
            Pipeline { st: self.st, _m: std::marker::PhantomData }
        }
        pub fn finish(self) -> (Sender<I>, Receiver<O>) {
            (Sender { st: self.st.clone(), _m: std::marker::PhantomData }, Receiver { st: self.st, _m: std::marker::PhantomData })
        }
    }
    impl<I: 'static> Sender<I> {
        pub fn send(&self, item: I) {
            let cell: *mut Option<Boxed> = Box::into_raw(Box::new(Some(Box::new(item) as Boxed)));
            let h = self.st.borrow().h;
            let st = unsafe { sys::ae_pipeline_send(h, cell as *mut c_void) };
            if st != sys::AE_OK { drop(unsafe { Box::from_raw(cell) }); }      // a failed stage: the item is not in flight
            ck(st)
        }
    }
    impl<O: 'static> Receiver<O> {
        pub fn recv(&self) -> O {
            let mut p = null_mut();
            let h = self.st.borrow().h;
            unsafe { ck(sys::ae_pipeline_recv(h, &mut p)) };
            let cell = unsafe { Box::from_raw(p as *mut Option<Boxed>) };
            *cell.expect("pipeline item").downcast::<O>().expect("pipeline output type")
        }
        pub fn in_flight(&self) -> usize { unsafe { sys::ae_pipeline_in_flight(self.st.borrow().h) } }
        /// per stage: processed, active time, rate, utilisation — what the reference's stage threads print (:93-107)
        pub fn report(&self, reset: bool) -> Vec<sys::ae_pipe_stage> {
            let h = self.st.borrow().h;
            let n = unsafe { sys::ae_pipeline_stages(h) };
            let mut st: Vec<sys::ae_pipe_stage> = Vec::with_capacity(n);
            unsafe { ck(sys::ae_pipeline_report(h, st.as_mut_ptr(), n, reset as c_int)); st.set_len(n) };
            st
        }
    }

    /// stream-ordered copies for the stages: no synchronisation; `host` is pinned and stays valid until the item is received
    pub unsafe fn upload_async(v: &mut DeviceVec, host: *const cf32, n: usize) { ck(sys::ae_vec_upload_async(v.h, host as *const _, n)) }
    pub unsafe fn download_async(v: &mut DeviceVec, host: *mut cf32, n: usize) { ck(sys::ae_vec_download_async(v.h, host as *mut _, n)) }
    pub unsafe fn upload_bits_async(b: &mut DeviceBits, host: *const u8, n: usize) { ck(sys::ae_bits_upload_async(b.h, host, n)) }
    pub unsafe fn download_bits_async(b: &mut DeviceBits, host: *mut u8, n: usize) { ck(sys::ae_bits_download_async(b.h, host, n)) }
}
