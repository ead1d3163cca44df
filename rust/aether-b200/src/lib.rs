//! Trait implementations of the reference crate on top of the C ABI (UNCOMPILED: no Rust toolchain in the build image).
use aether_primitives::{cf32, fft::{Fft, Scale}, vecops::VecOps};
use aether_b200_sys as sys;

fn ck(st: sys::ae_status) {
    if st != sys::AE_OK {
        // the reference panics (assert_eq!/unwrap); keep that contract, with its message text
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::ae_last_error_string()) };
        panic!("{}", msg.to_string_lossy());
    }
}
fn scale_args(s: Scale) -> (i32, f32) {
    match s { Scale::None => (0, 1.0), Scale::SN => (1, 1.0), Scale::N => (2, 1.0), Scale::X(x) => (3, x) }
}

/// Device-resident `Vec<cf32>`.  Send, not Sync: one caller at a time, like `&mut [cf32]`.
pub struct DeviceVec { h: *mut sys::ae_vec }
unsafe impl Send for DeviceVec {}
impl DeviceVec {
    pub fn from_slice(v: &[cf32]) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { ck(sys::ae_vec_alloc(v.len(), v.len(), &mut h)); ck(sys::ae_vec_upload(h, v.as_ptr() as *const _, v.len())); }
        DeviceVec { h }
    }
    pub fn to_vec(&mut self) -> Vec<cf32> {
        let n = unsafe { sys::ae_vec_len(self.h) };
        let mut out = vec![cf32::default(); n];
        unsafe { ck(sys::ae_vec_download(self.h, out.as_mut_ptr() as *mut _, n)); }
        out
    }
}
impl Drop for DeviceVec { fn drop(&mut self) { unsafe { sys::ae_vec_free(self.h); } } }

// The trait's `other: impl AsRef<[cf32]>` operands are host slices in the crate; on the device the
// operand is another DeviceVec, so the impl is for `&DeviceVec` operands via a small extension
// trait with the same method names.  User code changes only its buffer type.
pub trait DeviceVecOps {
    fn vec_scale(&mut self, scale: f32) -> &mut Self;
    fn vec_mul(&mut self, other: &DeviceVec) -> &mut Self;
    fn vec_div(&mut self, other: &DeviceVec) -> &mut Self;
    fn vec_conj(&mut self) -> &mut Self;
    fn vec_mirror(&mut self) -> &mut Self;
    fn vec_clone(&mut self, other: &DeviceVec) -> &mut Self;
    fn vec_zero(&mut self) -> &mut Self;
    fn vec_mutate(&mut self, f: impl FnMut(&mut cf32)) -> &mut Self;
    fn vec_add(&mut self, other: &DeviceVec) -> &mut Self;
    fn vec_sub(&mut self, other: &DeviceVec) -> &mut Self;
    fn vec_fft(&mut self, scale: Scale) -> &mut Self;
    fn vec_ifft(&mut self, scale: Scale) -> &mut Self;
    fn vec_rfft(&mut self, fft: &mut CudaFft, scale: Scale) -> &mut Self;
    fn vec_rifft(&mut self, fft: &mut CudaFft, scale: Scale) -> &mut Self;
}
impl DeviceVecOps for DeviceVec {
    fn vec_scale(&mut self, s: f32) -> &mut Self { unsafe { ck(sys::ae_vec_scale(self.h, s)) }; self }      // src/vecops.rs:94-97
    fn vec_mul(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_mul(self.h, o.h)) }; self } // :99-112 (AE_ELEN -> "Vectors must have same length")
    fn vec_div(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_div(self.h, o.h)) }; self }
    fn vec_conj(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_conj(self.h)) }; self }
    fn vec_mirror(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_mirror(self.h)) }; self }
    fn vec_clone(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_clone(self.h, o.h)) }; self }
    fn vec_zero(&mut self) -> &mut Self { unsafe { ck(sys::ae_vec_zero(self.h)) }; self }
    fn vec_mutate(&mut self, mut f: impl FnMut(&mut cf32)) -> &mut Self {
        extern "C" fn tramp<F: FnMut(&mut cf32)>(e: *mut sys::ae_cf32, u: *mut std::os::raw::c_void) {
            unsafe { (*(u as *mut F))(&mut *(e as *mut cf32)) }
        }
        fn call<F: FnMut(&mut cf32)>(h: *mut sys::ae_vec, f: &mut F) {
            unsafe { ck(sys::ae_vec_mutate(h, tramp::<F>, f as *mut F as *mut _)) }
        }
        call(self.h, &mut f); self                                                                            // host round trip (slow path)
    }
    fn vec_add(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_add(self.h, o.h)) }; self }
    fn vec_sub(&mut self, o: &DeviceVec) -> &mut Self { unsafe { ck(sys::ae_vec_sub(self.h, o.h)) }; self }
    fn vec_fft(&mut self, s: Scale) -> &mut Self { let (k, x) = scale_args(s); unsafe { ck(sys::ae_vec_fft(self.h, k, x, 0)) }; self }
    fn vec_ifft(&mut self, s: Scale) -> &mut Self { let (k, x) = scale_args(s); unsafe { ck(sys::ae_vec_ifft(self.h, k, x, 0)) }; self }
    fn vec_rfft(&mut self, fft: &mut CudaFft, s: Scale) -> &mut Self { fft.ifwd_dev(self, s); self }
    fn vec_rifft(&mut self, fft: &mut CudaFft, s: Scale) -> &mut Self { fft.ibwd_dev(self, s); self }
}

/// `Cfft`-named alias under feature `fft_b200`, so `use aether_primitives::fft::Cfft` call sites compile unchanged.
pub struct CudaFft { h: *mut sys::ae_fft, len: usize, io: DeviceVec }
unsafe impl Send for CudaFft {}
impl CudaFft {
    pub fn with_len(len: usize) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { ck(sys::ae_fft_create(len, &mut h)); }
        CudaFft { h, len, io: DeviceVec::from_slice(&vec![cf32::default(); len]) }
    }
    pub fn ifwd_dev(&mut self, v: &mut DeviceVec, s: Scale) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, 0, v.h, std::ptr::null_mut(), k, x, 1)) } }
    pub fn ibwd_dev(&mut self, v: &mut DeviceVec, s: Scale) { let (k, x) = scale_args(s); unsafe { ck(sys::ae_fft_exec(self.h, 1, v.h, std::ptr::null_mut(), k, x, 1)) } }
}
/// The crate's own `Fft` trait on HOST slices: upload, transform, download — a drop-in for
/// `impl Fft for Cfft` (src/fft.rs:161-235) wherever user code holds `&[cf32]`.
impl Fft for CudaFft {
    fn fwd(&mut self, input: &[cf32], output: &mut [cf32], s: Scale) {
        assert_eq!(self.len, input.len(), "Input and FFT must be the same length");
        unsafe { ck(sys::ae_vec_upload(self.io.h, input.as_ptr() as *const _, input.len())); }
        let (k, x) = scale_args(s);
        unsafe { ck(sys::ae_fft_exec(self.h, 0, self.io.h, std::ptr::null_mut(), k, x, 1));
                 ck(sys::ae_vec_download(self.io.h, output.as_mut_ptr() as *mut _, output.len())); }
    }
    fn bwd(&mut self, input: &[cf32], output: &mut [cf32], s: Scale) { /* same with dir = 1 */ unimplemented!() }
    fn ifwd(&mut self, input: &mut [cf32], s: Scale) { let tmp = input.to_vec(); self.fwd(&tmp, input, s) }
    fn ibwd(&mut self, input: &mut [cf32], s: Scale) { let tmp = input.to_vec(); self.bwd(&tmp, input, s) }
    fn tfwd(&mut self, _input: &[cf32], _s: Scale) -> &[cf32] { unimplemented!("host copy of ae_fft_exec_tmp's view") }
    fn tbwd(&mut self, _input: &[cf32], _s: Scale) -> &[cf32] { unimplemented!() }
    fn len(&self) -> usize { self.len }
}
impl Drop for CudaFft { fn drop(&mut self) { unsafe { sys::ae_fft_destroy(self.h); } } }
