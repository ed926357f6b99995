// Replacement bodies for halo2_proofs/src/arithmetic.rs and halo2_proofs/src/poly/domain.rs
// @ v2023_01_20.  SOURCE ONLY: no Rust toolchain exists in this image.
//
// The bn256 types are plain `[u64; 4]` Montgomery limbs (`Fr`, `Fq`), `G1Affine { x, y }` and
// `G1 { x, y, z }`; the const assertions below make the layout assumption explicit.
use std::any::TypeId;
use std::mem::{size_of, transmute_copy};

use b200zk_sys as ffi;
use halo2curves::bn256::{Fr, G1Affine, G1};

const _: () = assert!(size_of::<Fr>() == 32 && size_of::<G1Affine>() == 64 && size_of::<G1>() == 96);

/// `arithmetic::best_multiexp` — signature unchanged.
pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve {
    assert_eq!(coeffs.len(), bases.len());
    if TypeId::of::<C>() == TypeId::of::<G1Affine>() {
        let mut out = [0u64; 12];
        ffi::check(unsafe {
            ffi::b200zk_msm_g1(coeffs.as_ptr() as *const u64, bases.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr())
        });
        // a plain slice carries no mirror flag: if mirrors are on, forget what this call uploaded
        unsafe { ffi::b200zk_mirror_invalidate(coeffs.as_ptr() as *const core::ffi::c_void, 0) };
        // SAFETY: C::Curve == G1 here; [u64; 12] is its in-memory representation.
        return unsafe { transmute_copy::<[u64; 12], C::Curve>(&out) };
    }
    upstream::best_multiexp(coeffs, bases) // unchanged generic CPU path for other curves
}

/// `arithmetic::best_fft` — signature unchanged.
pub fn best_fft<G: Group>(a: &mut [G], omega: G::Scalar, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    if TypeId::of::<G>() == TypeId::of::<Fr>() {
        ffi::check(unsafe { ffi::b200zk_ntt(a.as_mut_ptr() as *mut u64, log_n, &omega as *const _ as *const u64) });
        unsafe { ffi::b200zk_mirror_invalidate(a.as_ptr() as *const core::ffi::c_void, 0) };   // `&mut [G]`: no flag to keep it coherent
        return;
    }
    upstream::best_fft(a, omega, log_n)
}

impl<G: Group> EvaluationDomain<G> {
    /// `EvaluationDomain::lagrange_to_coeff` — ifft + divisor in one device call.
    pub fn lagrange_to_coeff(&self, mut a: Polynomial<G, LagrangeCoeff>) -> Polynomial<G, Coeff> {
        assert_eq!(a.values.len(), 1 << self.k);
        if TypeId::of::<G>() == TypeId::of::<Fr>() {
            // `a.values` is touched through the field, not through DerefMut: a mirror left by commit_lagrange(&a)
            // stays valid, the transform runs in it and leaves it equal to what it writes back (poly_patch.rs)
            a.mark_mirrored();
            ffi::check(unsafe {
                ffi::b200zk_intt(a.values.as_mut_ptr() as *mut u64, self.k,
                                 &self.omega_inv as *const _ as *const u64, &self.ifft_divisor as *const _ as *const u64)
            });
        } else {
            Self::ifft(&mut a.values, self.omega_inv, self.k, self.ifft_divisor);
        }
        a.rebase()
    }

    /// `EvaluationDomain::coeff_to_extended` — zeta shift, zero padding and NTT fused.
    pub fn coeff_to_extended(&self, a: Polynomial<G, Coeff>) -> Polynomial<G, ExtendedLagrangeCoeff> {
        assert_eq!(a.values.len(), 1 << self.k);
        if TypeId::of::<G>() == TypeId::of::<Fr>() {
            let mut out = Polynomial::from_values(vec![G::group_zero(); self.extended_len()]);
            a.mark_mirrored();
            out.mark_mirrored();          // the library keeps the extended column for evaluate_h
            ffi::check(unsafe {
                ffi::b200zk_coeff_to_extended(a.values.as_ptr() as *const u64, self.k, out.values.as_mut_ptr() as *mut u64,
                                              self.extended_k, &self.extended_omega as *const _ as *const u64,
                                              &self.g_coset as *const _ as *const u64)
            });
            return out;
        }
        upstream::coeff_to_extended(self, a)
    }

    /// `EvaluationDomain::extended_to_coeff` — inverse NTT, un-shift and truncation fused.
    pub fn extended_to_coeff(&self, a: Polynomial<G, ExtendedLagrangeCoeff>) -> Vec<G> {
        assert_eq!(a.values.len(), self.extended_len());
        if TypeId::of::<G>() == TypeId::of::<Fr>() {
            let keep = (self.n * self.quotient_poly_degree) as usize;
            let mut out = vec![G::group_zero(); keep];
            ffi::check(unsafe {
                ffi::b200zk_extended_to_coeff(a.values.as_ptr() as *const u64, self.extended_k,
                                              &self.extended_omega_inv as *const _ as *const u64,
                                              &self.extended_ifft_divisor as *const _ as *const u64,
                                              &self.g_coset as *const _ as *const u64, out.as_mut_ptr() as *mut u64, keep)
            });
            return out;
        }
        upstream::extended_to_coeff(self, a)
    }

    /// `EvaluationDomain::divide_by_vanishing_poly`.
    pub fn divide_by_vanishing_poly(&self, mut a: Polynomial<G, ExtendedLagrangeCoeff>) -> Polynomial<G, ExtendedLagrangeCoeff> {
        assert_eq!(a.values.len(), self.extended_len());
        if TypeId::of::<G>() == TypeId::of::<Fr>() {
            ffi::check(unsafe {
                ffi::b200zk_divide_by_vanishing(a.values.as_mut_ptr() as *mut u64, self.extended_k,
                                                self.t_evaluations.as_ptr() as *const u64, self.t_evaluations.len() as u32)
            });
            return a;
        }
        upstream::divide_by_vanishing_poly(self, a)
    }
}

// poly/kzg/commitment.rs (`ParamsKZG::{commit, commit_lagrange, Drop}` and the handle fields): commitment_patch.rs
// plonk/evaluation.rs (`Evaluator::evaluate_h`, the flattened `GraphEvaluator`, the device-resident key): evaluation_patch.rs
// poly.rs (`Polynomial` with the mirror flag that keeps b200zk_mirror_* coherent): poly_patch.rs

// ---- the loops either side of the hot path (SURVEY.md section 8 f) ---------------------------

/// `arithmetic::eval_polynomial` — signature unchanged; bn256::Fr goes to the device.
pub fn eval_polynomial<F: Field>(poly: &[F], point: F) -> F {
    if TypeId::of::<F>() == TypeId::of::<Fr>() && poly.len() >= (1 << 12) {
        let mut h = 0u64;
        let mut out = [0u64; 4];
        unsafe {
            ffi::check(ffi::b200zk_dev_alloc(poly.len(), &mut h));
            ffi::check(ffi::b200zk_dev_upload(h, 0, poly.as_ptr() as *const u64, poly.len()));
            ffi::check(ffi::b200zk_eval_polynomial_dev(ffi::b200zk_dev_ptr(h), poly.len(), 1, poly.len(),
                                                       &point as *const _ as *const u64, out.as_mut_ptr(), std::ptr::null_mut()));
            ffi::check(ffi::b200zk_dev_free(h));
            return transmute_copy::<[u64; 4], F>(&out);
        }
    }
    upstream::eval_polynomial(poly, point)
}

/// `arithmetic::kate_division` — a(X) / (X - b), signature unchanged.
pub fn kate_division<'a, F: Field, I: IntoIterator<Item = &'a F>>(a: I, b: F) -> Vec<F>
where
    I::IntoIter: DoubleEndedIterator + ExactSizeIterator,
{
    let a: Vec<F> = a.into_iter().copied().collect();
    if TypeId::of::<F>() == TypeId::of::<Fr>() && a.len() >= (1 << 12) {
        let mut q = vec![F::zero(); a.len() - 1];
        let (mut ha, mut hq) = (0u64, 0u64);
        unsafe {
            ffi::check(ffi::b200zk_dev_alloc(a.len(), &mut ha));
            ffi::check(ffi::b200zk_dev_alloc(q.len(), &mut hq));
            ffi::check(ffi::b200zk_dev_upload(ha, 0, a.as_ptr() as *const u64, a.len()));
            ffi::check(ffi::b200zk_kate_division_dev(ffi::b200zk_dev_ptr(ha), a.len(), &b as *const _ as *const u64,
                                                     ffi::b200zk_dev_ptr(hq), std::ptr::null_mut()));
            ffi::check(ffi::b200zk_dev_download(hq, 0, q.as_mut_ptr() as *mut u64, q.len()));
            ffi::check(ffi::b200zk_dev_free(ha));
            ffi::check(ffi::b200zk_dev_free(hq));
        }
        return q;
    }
    upstream::kate_division(&a, b)
}

/// The `modified_values.batch_invert()` / `lookup_product.iter_mut().batch_invert()` calls of the
/// permutation and lookup provers (ff::BatchInvert), specialised for slices of bn256::Fr.
pub fn batch_invert_fr(values: &mut [Fr]) {
    ffi::check(unsafe { ffi::b200zk_batch_invert(values.as_mut_ptr() as *mut u64, values.len()) });
}

/// `G1Affine::from(point).to_bytes()` for a batch of commitments on their way into the transcript.
pub fn g1_to_bytes(points: &[G1]) -> Vec<[u8; 32]> {
    let mut out = vec![[0u8; 32]; points.len()];
    ffi::check(unsafe { ffi::b200zk_g1_to_bytes(points.as_ptr() as *const u64, points.len(), out.as_mut_ptr() as *mut u8) });
    out
}

// ---- page-locked host buffers ------------------------------------------------------------------
/// RAII guard around `b200zk_host_register`: page-locks the backing store of a `Vec<Fr>` /
/// `Polynomial` for as long as the guard lives, so the host-pointer entry points transfer at the
/// full PCIe rate and the MSM upload / NTT transfer pipelines can overlap their copies (on a B200
/// box: 72 ms instead of 170-179 ms for a 2^24 commit + transform from a pageable buffer).
/// `create_proof` takes one guard per long-lived polynomial vector (advice, permutation products,
/// lookup columns, h pieces); `ParamsKZG::read` takes one around the point buffer it registers.
pub struct PageLocked<'a, T> {
    slice: &'a [T],
}

impl<'a, T> PageLocked<'a, T> {
    pub fn new(slice: &'a [T]) -> Self {
        if !slice.is_empty() {
            ffi::check(unsafe {
                ffi::b200zk_host_register(slice.as_ptr() as *mut core::ffi::c_void, core::mem::size_of_val(slice))
            });
        }
        PageLocked { slice }
    }
}

impl<'a, T> Drop for PageLocked<'a, T> {
    fn drop(&mut self) {
        if !self.slice.is_empty() {
            ffi::check(unsafe { ffi::b200zk_host_unregister(self.slice.as_ptr() as *mut core::ffi::c_void) });
        }
    }
}
