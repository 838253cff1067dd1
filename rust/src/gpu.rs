//! Process-wide GPU state for the `cuda` feature: the context (one B200 or all of a node's), the cache
//! of base vectors that are resident in HBM, point decoding and the layout self-test.
//!
//! The reference hands `multiexp` an `(Arc<Vec<G::Affine>>, usize)` source per call
//! (src/groth16/mod.rs:438-477); a CRS vector is uploaded (and its window tables built) the first
//! time its `Arc` is seen and found again by pointer + length afterwards, which is what makes the
//! reference's call shape affordable: ~2.3 GB of points at 2^22 constraints are never re-sent.
//! Not compiled in the authoring image (no Rust toolchain there): see rust/README.md.
use std::any::{Any, TypeId};
use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::c_int;
use std::sync::{Arc, Mutex};

use bls12_381::{G1Affine, G1Projective, G2Affine, G2Projective, Scalar};
use ff::{FieldBits, PrimeField};
use lazy_static::lazy_static;

use crate::gpu_ffi as ffi;
use crate::SynthesisError;

/// Either one device (`bmpc_ctx`) or all devices named in BELLMAN_B200_DEVICES (`bmpc_multi`).
pub enum Ctx {
    Single(*mut ffi::bmpc_ctx),
    Multi(*mut ffi::bmpc_multi),
}
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

/// A base vector resident in HBM.
#[derive(Clone, Copy)]
pub enum Resident {
    Single(*const ffi::bmpc_bases),
    Multi(*const ffi::bmpc_multi_bases),
}
unsafe impl Send for Resident {}
unsafe impl Sync for Resident {}

pub struct Gpu {
    pub ctx: Ctx,
    /// (Arc data pointer, len, group) -> resident handle.  Entries live as long as the process: a CRS
    /// is loaded once per prover process in the reference's usage (`Parameters::read`).
    cache: Mutex<HashMap<(usize, usize, c_int), Resident>>,
}

lazy_static! {
    pub static ref GPU: Gpu = Gpu::new().expect("bellman-b200: no usable CUDA device (there is no CPU fallback)");
}

impl Gpu {
    fn new() -> Result<Gpu, SynthesisError> {
        layout_self_test();
        let devices: Vec<c_int> = std::env::var("BELLMAN_B200_DEVICES")
            .ok()
            .map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect())
            .unwrap_or_else(|| vec![0]);
        let ctx = if devices.len() > 1 {
            let mut m = std::ptr::null_mut();
            ffi::to_result(unsafe { ffi::bmpc_multi_create(devices.as_ptr(), devices.len() as c_int, &mut m) },
                           "bmpc_multi_create")?;
            Ctx::Multi(m)
        } else {
            let mut c = std::ptr::null_mut();
            ffi::to_result(unsafe { ffi::bmpc_ctx_create(devices[0], &mut c) }, "bmpc_ctx_create")?;
            Ctx::Single(c)
        };
        Ok(Gpu { ctx, cache: Mutex::new(HashMap::new()) })
    }

    pub fn last_error(&self) -> String {
        let p = match self.ctx {
            Ctx::Single(c) => unsafe { ffi::bmpc_last_error(c) },
            Ctx::Multi(m) => unsafe { ffi::bmpc_multi_last_error(m) },
        };
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }

    pub fn check(&self, st: c_int) -> Result<(), SynthesisError> {
        if st == ffi::BMPC_OK { Ok(()) } else { ffi::to_result(st, &self.last_error()) }
    }

    /// Registers `points` (ZCash uncompressed encoding, decoded on the device) once per Arc and
    /// builds the window tables; later calls find the handle by pointer.
    fn resident(&self, key: (usize, usize, c_int), encode: impl FnOnce() -> Vec<u8>) -> Result<Resident, SynthesisError> {
        if let Some(r) = self.cache.lock().unwrap().get(&key) {
            return Ok(*r);
        }
        let bytes = encode();
        let (n, group) = (key.1, key.2);
        let stride = if group == ffi::BMPC_G1 { 96 } else { 192 };
        let r = match self.ctx {
            Ctx::Single(c) => {
                let mut h = std::ptr::null_mut();
                self.check(unsafe {
                    ffi::bmpc_bases_register(c, group, bytes.as_ptr() as *const _, n, stride,
                                             ffi::BMPC_FORM_UNCOMPRESSED_BE, &mut h)
                })?;
                self.check(unsafe { ffi::bmpc_bases_precompute(c, h, 0) })?;
                Resident::Single(h)
            }
            Ctx::Multi(m) => {
                let mut h = std::ptr::null_mut();
                self.check(unsafe {
                    ffi::bmpc_multi_bases_register(m, group, bytes.as_ptr() as *const _, n, stride,
                                                   ffi::BMPC_FORM_UNCOMPRESSED_BE, &mut h)
                })?;
                self.check(unsafe { ffi::bmpc_multi_bases_precompute(m, h, 0) })?;
                Resident::Multi(h)
            }
        };
        self.cache.lock().unwrap().insert(key, r);
        Ok(r)
    }

    pub fn resident_g1(&self, v: &Arc<Vec<G1Affine>>) -> Result<Resident, SynthesisError> {
        self.resident((v.as_ptr() as usize, v.len(), ffi::BMPC_G1), || {
            let mut out = Vec::with_capacity(v.len() * 96);
            for p in v.iter() { out.extend_from_slice(&p.to_uncompressed()); }
            out
        })
    }

    pub fn resident_g2(&self, v: &Arc<Vec<G2Affine>>) -> Result<Resident, SynthesisError> {
        self.resident((v.as_ptr() as usize, v.len(), ffi::BMPC_G2), || {
            let mut out = Vec::with_capacity(v.len() * 192);
            for p in v.iter() { out.extend_from_slice(&p.to_uncompressed()); }
            out
        })
    }
}

/// `S` is generic in the reference (`SourceBuilder<G::Affine>`); the only builder the crate ever
/// passes is `(Arc<Vec<Affine>>, usize)` (src/multiexp.rs:45-51, groth16/mod.rs:438-477).  Returns
/// the resident handle, the start index and the group, or None for any other engine / builder.
pub fn resident_bases<G: 'static, S: Any>(bases: &S) -> Option<Result<(Resident, usize, c_int), SynthesisError>> {
    let any = bases as &dyn Any;
    if TypeId::of::<G>() == TypeId::of::<G1Projective>() {
        let (v, off) = any.downcast_ref::<(Arc<Vec<G1Affine>>, usize)>()?;
        return Some(GPU.resident_g1(v).map(|r| (r, *off, ffi::BMPC_G1)));
    }
    if TypeId::of::<G>() == TypeId::of::<G2Projective>() {
        let (v, off) = any.downcast_ref::<(Arc<Vec<G2Affine>>, usize)>()?;
        return Some(GPU.resident_g2(v).map(|r| (r, *off, ffi::BMPC_G2)));
    }
    None
}

/// 96 / 192 uncompressed bytes -> G (the library has already produced a canonical affine point of
/// the prime-order subgroup, so the unchecked decoder is the right one).
pub fn decode_point<G: 'static + Copy>(group: c_int, bytes: &[u8; 192]) -> G {
    if group == ffi::BMPC_G1 {
        let mut b = [0u8; 96];
        b.copy_from_slice(&bytes[..96]);
        let p: G1Projective = G1Affine::from_uncompressed_unchecked(&b).unwrap().into();
        *(&p as &dyn Any).downcast_ref::<G>().expect("group mismatch")
    } else {
        let p: G2Projective = G2Affine::from_uncompressed_unchecked(bytes).unwrap().into();
        *(&p as &dyn Any).downcast_ref::<G>().expect("group mismatch")
    }
}

/// The zero-copy facts the shim relies on, checked once at start-up:
/// `FieldBits<[u64; 4]>` and `bls12_381::Scalar` are 32 bytes of four little-endian u64 limbs; the
/// latter in Montgomery form (one = R mod r), which is what `bmpc_ntt` / `bmpc_create_proof` read.
pub fn layout_self_test() {
    use std::mem::size_of;
    assert_eq!(size_of::<FieldBits<[u64; 4]>>(), 32, "FieldBits<[u64;4]> is not a transparent [u64;4]");
    assert_eq!(size_of::<Scalar>(), 32, "bls12_381::Scalar is not [u64;4]");
    assert_eq!(size_of::<crate::domain::Scalar<Scalar>>(), 32, "domain::Scalar is not transparent");
    assert_eq!(size_of::<usize>(), 8, "density words are read as u64");
    // R mod r for BLS12-381's scalar field (bls12_381 0.6.0 src/scalar.rs `R`)
    const R: [u64; 4] = [0x0000_0001_ffff_fffe, 0x5884_b7fa_0003_4802, 0x998c_4fef_ecbc_4ff5, 0x1824_b159_acc5_056f];
    let one = Scalar::one();
    let limbs: [u64; 4] = unsafe { std::mem::transmute(one) };
    assert_eq!(limbs, R, "bls12_381::Scalar is not held as Montgomery limbs");
    // canonical form of the exponents: to_le_bits of 1 is the integer 1
    use ff::PrimeFieldBits;
    let bits = Scalar::one().to_le_bits();
    let raw: [u64; 4] = unsafe { std::mem::transmute(bits) };
    assert_eq!(raw, [1, 0, 0, 0], "FieldBits is not the canonical little-endian integer");
    let _ = Scalar::NUM_BITS;
}
