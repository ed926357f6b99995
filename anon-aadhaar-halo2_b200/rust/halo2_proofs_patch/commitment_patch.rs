// Replacement bodies for halo2_proofs/src/poly/kzg/commitment.rs @ v2023_01_20 (`ParamsKZG<Bn256>`).
// SOURCE ONLY (no Rust toolchain in this image).
//
// `g` and `g_lagrange` are fixed for the life of the params, so they live on the device (points plus
// the per-window table, b200zk_bases_register) and `commit` / `commit_lagrange` pass only the
// scalars.  Signatures unchanged; the struct gains two private handle fields and a Drop impl.
use b200zk_sys as ffi;
use halo2curves::bn256::{Bn256, Fr, G1Affine, G1, G2Affine};

pub struct ParamsKZG<E: Engine> {
    pub(crate) k: u32,
    pub(crate) n: u64,
    pub(crate) g: Vec<E::G1Affine>,
    pub(crate) g_lagrange: Vec<E::G1Affine>,
    pub(crate) g2: E::G2Affine,
    pub(crate) s_g2: E::G2Affine,
    // added: device registrations of `g` and `g_lagrange` (0 = not registered, non-bn256 engines)
    g_handle: u64,
    g_lagrange_handle: u64,
}

fn register(bases: &[G1Affine]) -> u64 {
    let mut h = 0u64;
    // the upload reads `bases` once; page-lock it for the duration (ParamsKZG::read of a 2^24 SRS: 1 GiB)
    let _pin = crate::arithmetic::PageLocked::new(bases);
    ffi::check(unsafe { ffi::b200zk_bases_register(bases.as_ptr() as *const u64, bases.len(), &mut h) });
    h
}

impl ParamsKZG<Bn256> {
    /// Called at the end of `setup`, `read`, `read_custom` and `downsize` (every constructor and
    /// every place that replaces `g` / `g_lagrange`).
    pub(crate) fn register_bases(&mut self) {
        self.evict_bases();
        self.g_handle = register(&self.g);
        self.g_lagrange_handle = register(&self.g_lagrange);
    }

    fn evict_bases(&mut self) {
        for h in [&mut self.g_handle, &mut self.g_lagrange_handle] {
            if *h != 0 {
                ffi::check(unsafe { ffi::b200zk_bases_evict(*h) });
                *h = 0;
            }
        }
    }

    fn msm(handle: u64, scalars: &[Fr]) -> G1 {
        let mut out = [0u64; 12];
        ffi::check(unsafe { ffi::b200zk_msm_g1_registered(handle, scalars.as_ptr() as *const u64, scalars.len(), out.as_mut_ptr()) });
        // SAFETY: [u64; 12] is the in-memory representation of bn256::G1 (x, y, z Montgomery limbs)
        unsafe { std::mem::transmute::<[u64; 12], G1>(out) }
    }
}

impl Drop for ParamsKZG<Bn256> {
    fn drop(&mut self) {
        self.evict_bases();
    }
}

impl<'params> Params<'params, G1Affine> for ParamsKZG<Bn256> {
    // k(), n(), downsize() (+ register_bases), empty_msm(), write(), read() (+ register_bases): as upstream

    /// `ParamsKZG::commit_lagrange(&self, poly, _) -> G1` — signature unchanged.
    fn commit_lagrange(&self, poly: &Polynomial<Fr, LagrangeCoeff>, _: Blind<Fr>) -> G1 {
        assert!(self.g_lagrange.len() >= poly.len());          // upstream: `assert!(bases.len() >= size)`
        poly.mark_mirrored();                                   // the next call on `poly` is lagrange_to_coeff
        Self::msm(self.g_lagrange_handle, &poly.values)
    }
}

impl<'params> ParamsProver<'params, G1Affine> for ParamsKZG<Bn256> {
    /// `ParamsKZG::commit(&self, poly, _) -> G1` — signature unchanged.
    fn commit(&self, poly: &Polynomial<Fr, Coeff>, _: Blind<Fr>) -> G1 {
        assert!(self.g.len() >= poly.len());
        poly.mark_mirrored();
        Self::msm(self.g_handle, &poly.values)
    }
}
