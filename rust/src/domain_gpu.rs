//! `EvaluationDomain::{fft, ifft, coset_fft, icoset_fft}` (src/domain.rs:80-125) for the one
//! coefficient type the prover uses, `Scalar<bls12_381::Scalar>`.  Each method starts with
//! `if domain_gpu::transform(&mut self.coeffs, self.exp, OP) { return; }`; the library fuses what the
//! reference does in extra sweeps (the `minv` loop :88-98, `distribute_powers` :101-113) into the
//! passes of the transform, so the four operations are one call each.  Results are the same field
//! elements, limb for limb.
//! Not compiled in the authoring image (no Rust toolchain there): see rust/README.md.
use std::any::TypeId;

use crate::gpu::{Ctx, GPU};
use crate::gpu_ffi as ffi;

/// true: `coeffs` (m = 2^exp Montgomery scalars) were transformed in place on the GPU.
/// false: not the BLS12-381 scalar type -- the caller runs the generic CPU body.
pub fn transform<T: 'static>(coeffs: &mut [T], exp: u32, op: std::os::raw::c_int) -> bool {
    if TypeId::of::<T>() != TypeId::of::<crate::domain::Scalar<bls12_381::Scalar>>() {
        return false;
    }
    assert_eq!(coeffs.len(), 1usize << exp);
    // a multi-device context transforms on its first device (a domain of 2^26 coefficients is 2 GB;
    // the four-step form over several GPUs is driven through the bmpc_ntt_batch_dev family)
    let ctx = match GPU.ctx {
        Ctx::Single(c) => c,
        Ctx::Multi(m) => unsafe { ffi::bmpc_multi_ctx(m, 0) },
    };
    let st = unsafe { ffi::bmpc_ntt(ctx, coeffs.as_mut_ptr() as *mut u64, exp, op) };
    // there is no CPU fallback for BLS12-381 with the `cuda` feature: a CUDA failure is fatal here
    // because the reference's fft methods return ()
    GPU.check(st).expect("bellman-b200: EvaluationDomain transform failed");
    true
}
