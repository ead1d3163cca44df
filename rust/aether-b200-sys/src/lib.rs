//! extern "C" declarations of include/aether_b200.h (UNCOMPILED: no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy)] pub struct ae_cf32 { pub re: f32, pub im: f32 }   // == num_complex::Complex32 (src/lib.rs:12)
#[repr(C)] pub struct ae_stats { pub bit_errors: u64, pub n_bits: u64, pub err_pow: f64, pub ref_pow: f64 }
pub enum ae_vec {} pub enum ae_bits {} pub enum ae_fft {} pub enum ae_fir {} pub enum ae_mod {} pub enum ae_awgn {} pub enum ae_chain {}
pub type ae_status = c_int;
pub const AE_OK: c_int = 0; pub const AE_ELEN: c_int = 1; pub const AE_EARG: c_int = 2;
pub const AE_ECUDA: c_int = 3; pub const AE_EOOM: c_int = 5; pub const AE_EIDX: c_int = 6;

extern "C" {
    pub fn ae_init(device: c_int) -> ae_status;
    pub fn ae_sync() -> ae_status;
    pub fn ae_last_error_string() -> *const c_char;
    // Vec<cf32>
    pub fn ae_vec_alloc(len: usize, capacity: usize, out: *mut *mut ae_vec) -> ae_status;
    pub fn ae_vec_view(parent: *mut ae_vec, offset: usize, len: usize, out: *mut *mut ae_vec) -> ae_status;
    pub fn ae_vec_free(v: *mut ae_vec) -> ae_status;
    pub fn ae_vec_len(v: *const ae_vec) -> usize;
    pub fn ae_vec_upload(v: *mut ae_vec, host: *const ae_cf32, n: usize) -> ae_status;
    pub fn ae_vec_download(v: *mut ae_vec, host: *mut ae_cf32, n: usize) -> ae_status;
    // VecOps (src/vecops.rs:39-89)
    pub fn ae_vec_scale(v: *mut ae_vec, s: f32) -> ae_status;
    pub fn ae_vec_mul(v: *mut ae_vec, o: *mut ae_vec) -> ae_status;
    pub fn ae_vec_div(v: *mut ae_vec, o: *mut ae_vec) -> ae_status;
    pub fn ae_vec_conj(v: *mut ae_vec) -> ae_status;
    pub fn ae_vec_mirror(v: *mut ae_vec) -> ae_status;
    pub fn ae_vec_clone(v: *mut ae_vec, o: *mut ae_vec) -> ae_status;
    pub fn ae_vec_zero(v: *mut ae_vec) -> ae_status;
    pub fn ae_vec_add(v: *mut ae_vec, o: *mut ae_vec) -> ae_status;
    pub fn ae_vec_sub(v: *mut ae_vec, o: *mut ae_vec) -> ae_status;
    pub fn ae_vec_mutate(v: *mut ae_vec, f: extern "C" fn(*mut ae_cf32, *mut c_void), user: *mut c_void) -> ae_status;
    pub fn ae_vec_fft(v: *mut ae_vec, scale_kind: c_int, x: f32, compat: c_int) -> ae_status;
    pub fn ae_vec_ifft(v: *mut ae_vec, scale_kind: c_int, x: f32, compat: c_int) -> ae_status;
    // Fft / Cfft (src/fft.rs:48-77, :134-235)
    pub fn ae_fft_create(len: usize, out: *mut *mut ae_fft) -> ae_status;
    pub fn ae_fft_destroy(f: *mut ae_fft) -> ae_status;
    pub fn ae_fft_len(f: *const ae_fft) -> usize;
    pub fn ae_fft_exec(f: *mut ae_fft, dir: c_int, input: *mut ae_vec, output: *mut ae_vec,
                       scale_kind: c_int, x: f32, howmany: usize) -> ae_status;
    pub fn ae_fft_exec_tmp(f: *mut ae_fft, dir: c_int, input: *mut ae_vec, scale_kind: c_int, x: f32,
                           howmany: usize, view: *mut *mut ae_vec) -> ae_status;
    // Modulation (src/modulation.rs:94-149), Awgn (src/noise.rs), sampling, sequence: see the header
    pub fn ae_mod_qpsk(out: *mut *mut ae_mod) -> ae_status;
    pub fn ae_mod_modulate(m: *mut ae_mod, bits: *mut ae_bits, out: *mut ae_vec) -> ae_status;
    pub fn ae_mod_demod(m: *mut ae_mod, symbols: *mut ae_vec, out: *mut ae_bits, compat: c_int) -> ae_status;
    pub fn ae_awgn_create(power: f32, seed: u64, out: *mut *mut ae_awgn) -> ae_status;
    pub fn ae_awgn_apply(g: *mut ae_awgn, signal: *mut ae_vec, compat: c_int) -> ae_status;
    pub fn ae_interpolate(src: *mut ae_vec, dst: *mut ae_vec, n_between: usize, compat: c_int) -> ae_status;
    pub fn ae_downsample(src: *mut ae_vec, dst: *mut ae_vec, strict: c_int) -> ae_status;
}
