//! The BLS12-381 branch of `multiexp` (src/multiexp.rs:254-281).  Called from the top of the
//! reference function, after its query-size assert; returns None for every other engine / source so
//! the generic CPU body (and with it the crate's DummyEngine tests) is untouched.
//!
//! Like the reference's `pool.compute(..)`, the call returns at once: `bmpc_multiexp_async` enqueues
//! the scalar upload and the kernels on a lane of the context and `wait()` blocks for the result,
//! so `create_proof`'s eight multiexps (prover.rs:233-307) are all in flight before the first wait.
//! Not compiled in the authoring image (no Rust toolchain there): see rust/README.md.
use std::sync::Arc;

use ff::{FieldBits, PrimeFieldBits};
use group::prime::PrimeCurve;

use crate::gpu::{self, Ctx, Resident, GPU};
use crate::gpu_ffi as ffi;
use crate::multicore::{Waiter, Worker};
use crate::multiexp::{DensityTracker, FullDensity, QueryDensity, SourceBuilder};
use crate::SynthesisError;

/// Density words as the library reads them (bitvec `BitVec<Lsb0, usize>` raw storage), or None for
/// FullDensity.  `D: AsRef<Q>` with Q = FullDensity | DensityTracker are the only two maps the crate
/// has (src/multiexp.rs:88-157).
fn raw_words<Q: 'static, D: AsRef<Q>>(d: &D) -> Option<Option<&[usize]>> {
    use std::any::Any;
    let q = d.as_ref() as &dyn Any;
    if q.is::<FullDensity>() {
        return Some(None);
    }
    q.downcast_ref::<DensityTracker>().map(|t| Some(t.raw_words()))
}

struct SendWaiter(*mut ffi::bmpc_waiter);
unsafe impl Send for SendWaiter {}

pub fn try_multiexp<Q, D, G, S>(
    pool: &Worker,
    bases: &S,
    density_map: &D,
    exponents: &Arc<Vec<FieldBits<<G::Scalar as PrimeFieldBits>::ReprBits>>>,
) -> Option<Waiter<Result<G, SynthesisError>>>
where
    for<'a> &'a Q: QueryDensity,
    Q: 'static,
    D: Send + Sync + 'static + Clone + AsRef<Q>,
    G: PrimeCurve,
    G::Scalar: PrimeFieldBits,
    S: SourceBuilder<<G as PrimeCurve>::Affine>,
{
    let (handle, offset, group) = match gpu::resident_bases::<G, S>(bases)? {
        Ok(t) => t,
        Err(e) => return Some(Waiter::done(Err(e))),
    };
    let words = raw_words::<Q, D>(density_map)?;
    let n = exponents.len();
    let sc = exponents.as_ptr() as *const u64; // FieldBits<[u64; 4]>: canonical little-endian limbs
    let (wp, wl) = match words {
        Some(w) => (w.as_ptr() as *const u64, n),
        None => (std::ptr::null(), 0),
    };
    match (&GPU.ctx, handle) {
        (Ctx::Single(ctx), Resident::Single(h)) => {
            let mut w = std::ptr::null_mut();
            let st = unsafe { ffi::bmpc_multiexp_async(*ctx, h, offset, sc, n, wp, wl, &mut w) };
            if let Err(e) = GPU.check(st) {
                return Some(Waiter::done(Err(e)));
            }
            // the exponents and the density map stay alive inside the closure until the result is in
            let (keep_e, keep_d, w) = (exponents.clone(), density_map.clone(), SendWaiter(w));
            Some(pool.compute(move || {
                let mut out = [0u8; 192];
                let st = unsafe { ffi::bmpc_waiter_wait(w.0, out.as_mut_ptr()) };
                drop((keep_e, keep_d));
                GPU.check(st).map(|_| gpu::decode_point::<G>(group, &out))
            }))
        }
        (Ctx::Multi(m), Resident::Multi(h)) => {
            // all devices of the node, still one call (bmpc_multi_multiexp is host-blocking: it runs
            // on a pool thread, as the reference's multiexp_inner does)
            let (keep_e, keep_d) = (exponents.clone(), density_map.clone());
            let (m, h) = (*m as usize, h as usize);
            Some(pool.compute(move || {
                let words = raw_words::<Q, D>(&keep_d).unwrap();
                let (wp, wl) = match words {
                    Some(w) => (w.as_ptr() as *const u64, keep_e.len()),
                    None => (std::ptr::null(), 0),
                };
                let mut out = [0u8; 192];
                let st = unsafe {
                    ffi::bmpc_multi_multiexp(m as *mut _, h as *const _, offset, keep_e.as_ptr() as *const u64,
                                             keep_e.len(), wp, wl, out.as_mut_ptr())
                };
                GPU.check(st).map(|_| gpu::decode_point::<G>(group, &out))
            }))
        }
        _ => unreachable!("resident handle and context kind always agree"),
    }
}
