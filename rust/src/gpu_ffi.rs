//! `extern "C"` declarations for libbellman_b200.so -- a line-for-line mirror of
//! `include/bellman_b200.h` (the header is the source of truth; `tests/test_abi.py` keeps the header,
//! the ctypes binding and the built library in agreement).  Not compiled in the authoring image (no
//! Rust toolchain there): see rust/README.md.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

macro_rules! opaque {
    ($($name:ident),*) => { $( #[repr(C)] pub struct $name { _private: [u8; 0] } )* };
}
opaque!(bmpc_ctx, bmpc_bases, bmpc_domain, bmpc_waiter, bmpc_multi, bmpc_multi_bases);

// bmpc_status: SynthesisError variants this path can produce (src/lib.rs:355-370)
pub const BMPC_OK: c_int = 0;
pub const BMPC_ERR_UNEXPECTED_IDENTITY: c_int = 1;
pub const BMPC_ERR_UNEXPECTED_EOF: c_int = 2;
pub const BMPC_ERR_DEGREE_TOO_LARGE: c_int = 3;
pub const BMPC_ERR_LENGTH_MISMATCH: c_int = 4;
pub const BMPC_ERR_CUDA: c_int = 5;
pub const BMPC_ERR_INVALID: c_int = 6;
pub const BMPC_ERR_INVALID_DATA: c_int = 7;

pub const BMPC_G1: c_int = 1;
pub const BMPC_G2: c_int = 2;
pub const BMPC_FORM_UNCOMPRESSED_BE: c_int = 0;
pub const BMPC_FORM_MONT_XY: c_int = 1;
pub const BMPC_FFT: c_int = 0;
pub const BMPC_IFFT: c_int = 1;
pub const BMPC_COSET_FFT: c_int = 2;
pub const BMPC_ICOSET_FFT: c_int = 3;
pub const BMPC_MSM_FLAG_EOF: u32 = 1;
pub const BMPC_MSM_FLAG_IDENT_ANY: u32 = 2;
pub const BMPC_MSM_FLAG_IDENT_TOP: u32 = 4;
pub const BMPC_PROOF_PARTIAL_BYTES: usize = 1920;

#[repr(C)]
pub struct bmpc_params {
    pub h: *const bmpc_bases,
    pub l: *const bmpc_bases,
    pub a: *const bmpc_bases,
    pub b_g1: *const bmpc_bases,
    pub b_g2: *const bmpc_bases,
    pub alpha_g1: [u8; 96],
    pub beta_g1: [u8; 96],
    pub beta_g2: [u8; 192],
    pub delta_g1: [u8; 96],
    pub delta_g2: [u8; 192],
}

#[repr(C)]
pub struct bmpc_multi_params {
    pub h: *const bmpc_multi_bases,
    pub l: *const bmpc_multi_bases,
    pub a: *const bmpc_multi_bases,
    pub b_g1: *const bmpc_multi_bases,
    pub b_g2: *const bmpc_multi_bases,
    pub alpha_g1: [u8; 96],
    pub beta_g1: [u8; 96],
    pub beta_g2: [u8; 192],
    pub delta_g1: [u8; 96],
    pub delta_g2: [u8; 192],
}

#[repr(C)]
pub struct bmpc_assignment {
    pub a: *const u64,
    pub b: *const u64,
    pub c: *const u64,
    pub num_constraints: usize,
    pub input_assignment: *const u64,
    pub num_inputs: usize,
    pub aux_assignment: *const u64,
    pub num_aux: usize,
    pub a_aux_density: *const u64,
    pub b_input_density: *const u64,
    pub b_aux_density: *const u64,
}

#[repr(C)]
pub struct bmpc_proof_shard {
    pub base_offset: [usize; 8],
    pub h_lo: usize,
    pub h_hi: usize,
    pub n_total: [usize; 8],
}

#[repr(C)]
pub struct bmpc_parameters {
    pub p: bmpc_params,
    pub gamma_g2: [u8; 192],
    pub ic: *mut bmpc_bases,
}

#[repr(C)]
pub struct bmpc_csr {
    pub row_ptr: *const u32,
    pub col: *const u32,
    pub coeff: *const u64,
    pub num_rows: usize,
    pub nnz: usize,
}

extern "C" {
    // ---- context (Worker::new, src/multicore.rs:25-27)
    pub fn bmpc_ctx_create(device: c_int, out: *mut *mut bmpc_ctx) -> c_int;
    pub fn bmpc_ctx_destroy(ctx: *mut bmpc_ctx);
    pub fn bmpc_last_error(ctx: *const bmpc_ctx) -> *const c_char;
    pub fn bmpc_ctx_set_tuning(ctx: *mut bmpc_ctx, msm_window_bits: c_int, ntt_max_deg: c_int) -> c_int;
    pub fn bmpc_ctx_reload_env(ctx: *mut bmpc_ctx) -> c_int;
    pub fn bmpc_ctx_launch_count(ctx: *const bmpc_ctx) -> u64;
    pub fn bmpc_ctx_profile(ctx: *mut bmpc_ctx, enable: c_int) -> c_int;
    pub fn bmpc_ctx_profile_read(ctx: *mut bmpc_ctx, which: c_int, ms_total: *mut f64, launches: *mut u64) -> c_int;

    // ---- bases (SourceBuilder for (Arc<Vec<G>>, usize), src/multiexp.rs:45-86)
    pub fn bmpc_bases_register(ctx: *mut bmpc_ctx, group: c_int, points: *const c_void, n: usize, stride: usize,
                               form: c_int, out: *mut *mut bmpc_bases) -> c_int;
    pub fn bmpc_bases_register_dev(ctx: *mut bmpc_ctx, group: c_int, d_points_mont: *const c_void, n: usize,
                                   out: *mut *mut bmpc_bases, stream: *mut c_void) -> c_int;
    pub fn bmpc_bases_precompute(ctx: *mut bmpc_ctx, b: *mut bmpc_bases, window_bits: c_int) -> c_int;
    pub fn bmpc_bases_len(b: *const bmpc_bases) -> usize;
    pub fn bmpc_bases_group(b: *const bmpc_bases) -> c_int;
    pub fn bmpc_bases_read(ctx: *mut bmpc_ctx, b: *const bmpc_bases, start: usize, count: usize, out: *mut u8) -> c_int;
    pub fn bmpc_bases_dev_ptr(b: *const bmpc_bases) -> *const c_void;
    pub fn bmpc_bases_free(ctx: *mut bmpc_ctx, b: *mut bmpc_bases);

    // ---- multiexp (src/multiexp.rs:159-281)
    pub fn bmpc_multiexp(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize, scalars: *const u64,
                         n: usize, density_words: *const u64, density_len: usize, out: *mut u8) -> c_int;
    pub fn bmpc_multiexp_dev(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize, d_scalars: *const u64,
                             n: usize, d_density_words: *const u64, density_len: usize, out: *mut u8,
                             stream: *mut c_void) -> c_int;
    pub fn bmpc_multiexp_partial_dev(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize,
                                     d_scalars: *const u64, n: usize, d_density_words: *const u64,
                                     density_len: usize, d_partial_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn bmpc_multiexp_shard_dev(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize,
                                   d_scalars: *const u64, n: usize, d_density_words: *const u64,
                                   density_len: usize, n_total: usize, d_partial_out: *mut c_void,
                                   flags_out: *mut u32, stream: *mut c_void) -> c_int;
    pub fn bmpc_msm_flags_status(flags_or: u32) -> c_int;
    pub fn bmpc_multiexp_shard_enqueue_dev(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize,
                                           d_scalars: *const u64, n: usize, d_density_words: *const u64,
                                           density_len: usize, n_total: usize, d_record_out: *mut c_void,
                                           stream: *mut c_void) -> c_int;
    pub fn bmpc_shard_record_bytes(group: c_int) -> usize;
    pub fn bmpc_fold_shard_records(ctx: *mut bmpc_ctx, group: c_int, d_records: *const c_void, count: usize,
                                   stride: usize, out: *mut u8, flags_or_out: *mut u32, stream: *mut c_void) -> c_int;
    pub fn bmpc_sum_partials(ctx: *mut bmpc_ctx, group: c_int, d_partials: *const c_void, count: usize,
                             out: *mut u8, stream: *mut c_void) -> c_int;
    pub fn bmpc_partial_bytes(group: c_int) -> usize;
    pub fn bmpc_msm_geometry(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, n: usize, window_bits: *mut u32,
                             windows: *mut u32, bucket_sets: *mut u32) -> c_int;
    pub fn bmpc_msm_accumulate_info(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, n: usize, info: *mut u32) -> c_int;
    // Waiter (src/multicore.rs:93-118)
    pub fn bmpc_multiexp_async(ctx: *mut bmpc_ctx, bases: *const bmpc_bases, base_offset: usize, scalars: *const u64,
                               n: usize, density_words: *const u64, density_len: usize,
                               out: *mut *mut bmpc_waiter) -> c_int;
    pub fn bmpc_waiter_wait(w: *mut bmpc_waiter, out: *mut u8) -> c_int;

    // ---- EvaluationDomain (src/domain.rs:21-189)
    pub fn bmpc_domain_from_coeffs(ctx: *mut bmpc_ctx, coeffs: *const u64, len: usize, out: *mut *mut bmpc_domain) -> c_int;
    pub fn bmpc_domain_from_coeffs_dev(ctx: *mut bmpc_ctx, d_coeffs: *const u64, len: usize,
                                       out: *mut *mut bmpc_domain, stream: *mut c_void) -> c_int;
    pub fn bmpc_domain_len(d: *const bmpc_domain) -> usize;
    pub fn bmpc_domain_exp(d: *const bmpc_domain) -> u32;
    pub fn bmpc_domain_into_coeffs(ctx: *mut bmpc_ctx, d: *const bmpc_domain, out: *mut u64) -> c_int;
    pub fn bmpc_domain_dev_ptr(d: *mut bmpc_domain) -> *mut u64;
    pub fn bmpc_domain_free(ctx: *mut bmpc_ctx, d: *mut bmpc_domain);
    pub fn bmpc_domain_transform(ctx: *mut bmpc_ctx, d: *mut bmpc_domain, op: c_int, stream: *mut c_void) -> c_int;
    pub fn bmpc_domain_distribute_powers(ctx: *mut bmpc_ctx, d: *mut bmpc_domain, g: *const u64, stream: *mut c_void) -> c_int;
    pub fn bmpc_domain_z(ctx: *mut bmpc_ctx, d: *const bmpc_domain, tau: *const u64, out: *mut u64) -> c_int;
    pub fn bmpc_domain_divide_by_z_on_coset(ctx: *mut bmpc_ctx, d: *mut bmpc_domain, stream: *mut c_void) -> c_int;
    pub fn bmpc_domain_mul_assign(ctx: *mut bmpc_ctx, d: *mut bmpc_domain, other: *const bmpc_domain, stream: *mut c_void) -> c_int;
    pub fn bmpc_domain_sub_assign(ctx: *mut bmpc_ctx, d: *mut bmpc_domain, other: *const bmpc_domain, stream: *mut c_void) -> c_int;
    pub fn bmpc_ntt_dev(ctx: *mut bmpc_ctx, d_coeffs: *mut u64, log_m: u32, op: c_int, stream: *mut c_void) -> c_int;
    pub fn bmpc_ntt(ctx: *mut bmpc_ctx, coeffs: *mut u64, log_m: u32, op: c_int) -> c_int;
    // pieces of the distributed four-step transform
    pub fn bmpc_ntt_batch_dev(ctx: *mut bmpc_ctx, d_coeffs: *mut u64, log_n: u32, batch: u32, inverse: c_int,
                              stream: *mut c_void) -> c_int;
    pub fn bmpc_fr_swap01_dev(ctx: *mut bmpc_ctx, d_in: *const u64, d_out: *mut u64, d0: u32, d1: u32, d2: u32,
                              stream: *mut c_void) -> c_int;
    pub fn bmpc_ntt_fourstep_twiddle_dev(ctx: *mut bmpc_ctx, d: *mut u64, rows: u32, cols: u32, row0: u32, log_m: u32,
                                         inverse: c_int, stream: *mut c_void) -> c_int;
    pub fn bmpc_fr_scale_pow_dev(ctx: *mut bmpc_ctx, d: *mut u64, n: usize, first: u32, log_m: u32, which: c_int,
                                 stream: *mut c_void) -> c_int;

    // ---- H polynomial and create_proof (src/groth16/prover.rs:176-350)
    pub fn bmpc_h_coefficients(ctx: *mut bmpc_ctx, a: *const u64, b: *const u64, c: *const u64, len: usize,
                               out: *mut u64, out_len: *mut usize) -> c_int;
    pub fn bmpc_h_coefficients_dev(ctx: *mut bmpc_ctx, d_a: *mut u64, d_b: *mut u64, d_c: *mut u64, log_m: u32,
                                   stream: *mut c_void) -> c_int;
    pub fn bmpc_h_coset_evals_dev(ctx: *mut bmpc_ctx, d_p: *mut u64, log_m: u32, host_src: *const u64, host_len: usize,
                                  stream: *mut c_void) -> c_int;
    pub fn bmpc_h_from_coset_evals_dev(ctx: *mut bmpc_ctx, d_a: *mut u64, d_b: *const u64, d_c: *const u64, log_m: u32,
                                       stream: *mut c_void) -> c_int;
    pub fn bmpc_fr_to_canonical_dev(ctx: *mut bmpc_ctx, d_vals: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn bmpc_create_proof(ctx: *mut bmpc_ctx, params: *const bmpc_params, asg: *const bmpc_assignment,
                             r: *const u64, s: *const u64, proof_out: *mut u8) -> c_int;
    pub fn bmpc_create_proof_partials(ctx: *mut bmpc_ctx, params: *const bmpc_params, asg: *const bmpc_assignment,
                                      shard: *const bmpc_proof_shard, partials_out: *mut u8, flags_out: *mut u32) -> c_int;
    pub fn bmpc_create_proof_finish(ctx: *mut bmpc_ctx, params: *const bmpc_params, partials_all: *const u8,
                                    world: usize, r: *const u64, s: *const u64, proof_out: *mut u8) -> c_int;

    // ---- all GPUs of a node behind one call
    pub fn bmpc_multi_create(devices: *const c_int, n: c_int, out: *mut *mut bmpc_multi) -> c_int;
    pub fn bmpc_multi_destroy(m: *mut bmpc_multi);
    pub fn bmpc_multi_size(m: *const bmpc_multi) -> c_int;
    pub fn bmpc_multi_ctx(m: *mut bmpc_multi, rank: c_int) -> *mut bmpc_ctx;
    pub fn bmpc_multi_last_error(m: *const bmpc_multi) -> *const c_char;
    pub fn bmpc_multi_bases_register(m: *mut bmpc_multi, group: c_int, points: *const c_void, n: usize,
                                     stride: usize, form: c_int, out: *mut *mut bmpc_multi_bases) -> c_int;
    pub fn bmpc_multi_bases_precompute(m: *mut bmpc_multi, b: *mut bmpc_multi_bases, window_bits: c_int) -> c_int;
    pub fn bmpc_multi_bases_len(b: *const bmpc_multi_bases) -> usize;
    pub fn bmpc_multi_bases_part(b: *const bmpc_multi_bases, rank: c_int, first: *mut usize) -> *const bmpc_bases;
    pub fn bmpc_multi_bases_free(m: *mut bmpc_multi, b: *mut bmpc_multi_bases);
    pub fn bmpc_multi_multiexp(m: *mut bmpc_multi, bases: *const bmpc_multi_bases, base_offset: usize,
                               scalars: *const u64, n: usize, density_words: *const u64, density_len: usize,
                               out: *mut u8) -> c_int;
    pub fn bmpc_multi_create_proof(m: *mut bmpc_multi, params: *const bmpc_multi_params, asg: *const bmpc_assignment,
                                   r: *const u64, s: *const u64, proof_out: *mut u8) -> c_int;

    // ---- Parameters wire format, constraint-system side, ceremony
    pub fn bmpc_params_read(ctx: *mut bmpc_ctx, data: *const u8, len: usize, checked: c_int, out: *mut bmpc_parameters) -> c_int;
    pub fn bmpc_params_write(ctx: *mut bmpc_ctx, p: *const bmpc_parameters, out: *mut u8, cap: usize, written: *mut usize) -> c_int;
    pub fn bmpc_params_free(ctx: *mut bmpc_ctx, p: *mut bmpc_parameters);
    pub fn bmpc_r1cs_eval(ctx: *mut bmpc_ctx, a: *const bmpc_csr, b: *const bmpc_csr, c: *const bmpc_csr,
                          num_inputs: usize, num_aux: usize, input_assignment: *const u64, aux_assignment: *const u64,
                          a_out: *mut u64, b_out: *mut u64, c_out: *mut u64, a_aux_density: *mut u64,
                          b_input_density: *mut u64, b_aux_density: *mut u64) -> c_int;
    pub fn bmpc_generate_parameters(ctx: *mut bmpc_ctx, at: *const bmpc_csr, bt: *const bmpc_csr, ct: *const bmpc_csr,
                                    num_inputs: usize, num_aux: usize, num_constraints: usize, g1: *const u8,
                                    g2: *const u8, alpha: *const u64, beta: *const u64, gamma: *const u64,
                                    delta: *const u64, tau: *const u64, out: *mut bmpc_parameters) -> c_int;
    pub fn bmpc_batch_scalar_mul(ctx: *mut bmpc_ctx, input: *const bmpc_bases, scalars: *const u64, per_element: c_int,
                                 out: *mut *mut bmpc_bases) -> c_int;
    pub fn bmpc_list_mul_matrix(ctx: *mut bmpc_ctx, list: *const bmpc_bases, row_ptr: *const u64, cols: *const u32,
                                coeffs: *const u64, n_rows: usize, out: *mut *mut bmpc_bases) -> c_int;
    pub fn bmpc_fixed_base_mul(ctx: *mut bmpc_ctx, group: c_int, base: *const u8, scalars: *const u64, n: usize,
                               scalars_on_device: c_int, out: *mut *mut bmpc_bases) -> c_int;
}

/// status -> the reference's error type (src/lib.rs:355-370).  `msg`: bmpc_last_error of the context.
pub fn to_result(st: c_int, msg: &str) -> Result<(), crate::SynthesisError> {
    use crate::SynthesisError::*;
    use std::io::{Error, ErrorKind};
    match st {
        BMPC_OK => Ok(()),
        BMPC_ERR_UNEXPECTED_IDENTITY => Err(UnexpectedIdentity),
        BMPC_ERR_UNEXPECTED_EOF => Err(IoError(Error::new(ErrorKind::UnexpectedEof, "expected more bases from source"))),
        BMPC_ERR_DEGREE_TOO_LARGE => Err(PolynomialDegreeTooLarge),
        BMPC_ERR_LENGTH_MISMATCH => panic!("assertion failed: length mismatch ({})", msg), // the reference asserts
        BMPC_ERR_INVALID_DATA => Err(IoError(Error::new(ErrorKind::InvalidData, msg.to_string()))),
        _ => Err(IoError(Error::new(ErrorKind::Other, format!("bellman-b200: {}", msg)))),
    }
}

#[cfg(test)]
mod layout {
    //! The header's structs as the C compiler lays them out on x86-64 / aarch64 (LP64).
    use super::*;
    use std::mem::{align_of, size_of};

    #[test]
    fn struct_sizes_match_the_header() {
        assert_eq!(size_of::<bmpc_params>(), 5 * 8 + 96 + 96 + 192 + 96 + 192);
        assert_eq!(size_of::<bmpc_multi_params>(), size_of::<bmpc_params>());
        assert_eq!(size_of::<bmpc_assignment>(), 11 * 8);
        assert_eq!(size_of::<bmpc_proof_shard>(), 18 * 8);
        assert_eq!(size_of::<bmpc_parameters>(), size_of::<bmpc_params>() + 192 + 8);
        assert_eq!(size_of::<bmpc_csr>(), 5 * 8);
        assert_eq!(align_of::<bmpc_params>(), 8);
    }
}
