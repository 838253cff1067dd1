//! `create_proof` after synthesis (src/groth16/prover.rs:206-350) as one library call: the seven
//! transforms of the H pipeline, the eight multiexps and the proof tail run on the GPU(s) and 192
//! proof bytes come back.  Inserted after `let vk = params.get_vk(..)?;`; returns None for any
//! engine other than bls12_381::Bls12 or a ParameterSource that is not `&Parameters`, so the
//! reference body stays for those.
//! Not compiled in the authoring image (no Rust toolchain there): see rust/README.md.
use std::any::{Any, TypeId};

use bls12_381::{Bls12, G1Affine, G2Affine};
use pairing::Engine;

use crate::gpu::{Ctx, Resident, GPU};
use crate::gpu_ffi as ffi;
use crate::groth16::{ParameterSource, Parameters, Proof, VerifyingKey};
use crate::SynthesisError;

use super::prover::ProvingAssignment; // same module in the reference: fields are reachable

fn fill_vk(vk: &VerifyingKey<Bls12>, alpha_g1: &mut [u8; 96], beta_g1: &mut [u8; 96], beta_g2: &mut [u8; 192],
           delta_g1: &mut [u8; 96], delta_g2: &mut [u8; 192]) {
    alpha_g1.copy_from_slice(&vk.alpha_g1.to_uncompressed());
    beta_g1.copy_from_slice(&vk.beta_g1.to_uncompressed());
    beta_g2.copy_from_slice(&vk.beta_g2.to_uncompressed());
    delta_g1.copy_from_slice(&vk.delta_g1.to_uncompressed());
    delta_g2.copy_from_slice(&vk.delta_g2.to_uncompressed());
}

pub fn try_create_proof<E: Engine + 'static, P: ParameterSource<E> + 'static>(
    params: &mut P,
    prover: &ProvingAssignment<E::Fr>,
    vk: &VerifyingKey<E>,
    r: E::Fr,
    s: E::Fr,
) -> Option<Result<Proof<E>, SynthesisError>> {
    if TypeId::of::<E>() != TypeId::of::<Bls12>() {
        return None;
    }
    let params = (params as &mut dyn Any).downcast_mut::<&Parameters<Bls12>>()?;
    let vk = (vk as &dyn Any).downcast_ref::<VerifyingKey<Bls12>>()?;
    let run = || -> Result<[u8; 192], SynthesisError> {
        // ParameterSource for &Parameters (groth16/mod.rs:438-477): the Arcs themselves
        let (h, l) = (GPU.resident_g1(&params.h)?, GPU.resident_g1(&params.l)?);
        let (a, b1) = (GPU.resident_g1(&params.a)?, GPU.resident_g1(&params.b_g1)?);
        let b2 = GPU.resident_g2(&params.b_g2)?;
        let asg = ffi::bmpc_assignment {
            a: prover.a.as_ptr() as *const u64,
            b: prover.b.as_ptr() as *const u64,
            c: prover.c.as_ptr() as *const u64,
            num_constraints: prover.a.len(),
            input_assignment: prover.input_assignment.as_ptr() as *const u64,
            num_inputs: prover.input_assignment.len(),
            aux_assignment: prover.aux_assignment.as_ptr() as *const u64,
            num_aux: prover.aux_assignment.len(),
            a_aux_density: prover.a_aux_density.raw_words().as_ptr() as *const u64,
            b_input_density: prover.b_input_density.raw_words().as_ptr() as *const u64,
            b_aux_density: prover.b_aux_density.raw_words().as_ptr() as *const u64,
        };
        let (rp, sp) = (&r as *const E::Fr as *const u64, &s as *const E::Fr as *const u64); // Montgomery limbs
        let mut proof = [0u8; 192];
        match (&GPU.ctx, h, l, a, b1, b2) {
            (Ctx::Single(ctx), Resident::Single(h), Resident::Single(l), Resident::Single(a), Resident::Single(b1),
             Resident::Single(b2)) => {
                let mut p = ffi::bmpc_params { h, l, a, b_g1: b1, b_g2: b2, alpha_g1: [0; 96], beta_g1: [0; 96],
                                               beta_g2: [0; 192], delta_g1: [0; 96], delta_g2: [0; 192] };
                fill_vk(vk, &mut p.alpha_g1, &mut p.beta_g1, &mut p.beta_g2, &mut p.delta_g1, &mut p.delta_g2);
                GPU.check(unsafe { ffi::bmpc_create_proof(*ctx, &p, &asg, rp, sp, proof.as_mut_ptr()) })?;
            }
            (Ctx::Multi(m), Resident::Multi(h), Resident::Multi(l), Resident::Multi(a), Resident::Multi(b1),
             Resident::Multi(b2)) => {
                let mut p = ffi::bmpc_multi_params { h, l, a, b_g1: b1, b_g2: b2, alpha_g1: [0; 96], beta_g1: [0; 96],
                                                     beta_g2: [0; 192], delta_g1: [0; 96], delta_g2: [0; 192] };
                fill_vk(vk, &mut p.alpha_g1, &mut p.beta_g1, &mut p.beta_g2, &mut p.delta_g1, &mut p.delta_g2);
                GPU.check(unsafe { ffi::bmpc_multi_create_proof(*m, &p, &asg, rp, sp, proof.as_mut_ptr()) })?;
            }
            _ => unreachable!("resident handles and context kind always agree"),
        }
        Ok(proof)
    };
    Some(run().and_then(|bytes| {
        // Proof::write's format (groth16/mod.rs:42-48): read it back into the engine's types
        let p = Proof::<Bls12>::read(&bytes[..]).map_err(SynthesisError::from)?;
        let any: Box<dyn Any> = Box::new(p);
        Ok(*any.downcast::<Proof<E>>().expect("E is Bls12"))
    }))
}

#[allow(dead_code)]
fn _types(_: G1Affine, _: G2Affine) {}
