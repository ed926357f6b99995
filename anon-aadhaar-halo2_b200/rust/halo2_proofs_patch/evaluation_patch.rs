// Replacement body for halo2_proofs/src/plonk/evaluation.rs `Evaluator::evaluate_h` @ v2023_01_20,
// for C = bn256::G1Affine.  SOURCE ONLY (no Rust toolchain in this image).
//
// `GraphEvaluator` (constants, rotations, calculations) is flattened into the `b200zk_graph`
// encoding once per proving key; the proving key's extended columns are uploaded once
// (`DeviceKey`, built lazily on the first proof and dropped with the key); per proof the advice /
// instance / lookup polynomials are staged (device-to-device when the library mirrors them),
// extended on the device, and the gate, permutation and lookup kernels fold their terms into one
// device column in upstream's order.  The signature and the returned `Polynomial` are unchanged.
use b200zk_sys as ffi;
use ffi::{b200zk_calc as Calc, b200zk_graph as Graph, b200zk_quotient_env as Env, b200zk_src as Src};
use halo2curves::bn256::{Fr, G1Affine};

fn limbs(x: &Fr) -> [u64; 4] {
    // SAFETY: bn256::Fr is four little-endian u64 Montgomery limbs
    unsafe { std::mem::transmute_copy::<Fr, [u64; 4]>(x) }
}

fn src(v: &ValueSource) -> Src {
    // discriminants as documented in include/b200zk.h (`b200zk_src`)
    match *v {
        ValueSource::Constant(i) => Src { kind: 0, a: i as u32, b: 0 },
        ValueSource::Intermediate(i) => Src { kind: 1, a: i as u32, b: 0 },
        ValueSource::Fixed(c, r) => Src { kind: 2, a: c as u32, b: r as u32 },
        ValueSource::Advice(c, r) => Src { kind: 3, a: c as u32, b: r as u32 },
        ValueSource::Instance(c, r) => Src { kind: 4, a: c as u32, b: r as u32 },
        ValueSource::Challenge(i) => Src { kind: 5, a: i as u32, b: 0 },
        ValueSource::Beta() => Src { kind: 6, a: 0, b: 0 },
        ValueSource::Gamma() => Src { kind: 7, a: 0, b: 0 },
        ValueSource::Theta() => Src { kind: 8, a: 0, b: 0 },
        ValueSource::Y() => Src { kind: 9, a: 0, b: 0 },
        ValueSource::PreviousValue() => Src { kind: 10, a: 0, b: 0 },
    }
}

/// `GraphEvaluator` in the flat form the device interprets (`b200zk_calc`: op 0 Add 1 Sub 2 Mul
/// 3 Square 4 Double 5 Negate 6 Horner 7 Store).
pub(crate) struct FlatGraph {
    constants: Vec<[u64; 4]>,
    rotations: Vec<i32>,
    calcs: Vec<Calc>,
    parts: Vec<Src>,
    n_intermediates: u32,
}

impl FlatGraph {
    pub(crate) fn new(g: &GraphEvaluator<G1Affine>) -> Self {
        let zero = Src { kind: 0, a: 0, b: 0 };
        let mut parts = Vec::new();
        let calcs = g.calculations.iter().map(|info| {
            let (op, x, y, off, len) = match &info.calculation {
                Calculation::Add(a, b) => (0, src(a), src(b), 0, 0),
                Calculation::Sub(a, b) => (1, src(a), src(b), 0, 0),
                Calculation::Mul(a, b) => (2, src(a), src(b), 0, 0),
                Calculation::Square(a) => (3, src(a), zero, 0, 0),
                Calculation::Double(a) => (4, src(a), zero, 0, 0),
                Calculation::Negate(a) => (5, src(a), zero, 0, 0),
                Calculation::Horner(start, ps, factor) => {
                    let off = parts.len() as u32;
                    parts.extend(ps.iter().map(src));
                    (6, src(start), src(factor), off, ps.len() as u32)
                }
                Calculation::Store(a) => (7, src(a), zero, 0, 0),
            };
            Calc { op, target: info.target as u32, x, y, parts_off: off, parts_len: len }
        }).collect();
        FlatGraph { constants: g.constants.iter().map(limbs).collect(), rotations: g.rotations.clone(), calcs, parts,
                    n_intermediates: g.num_intermediates as u32 }
    }

    fn as_c(&self) -> Graph {
        Graph { constants: self.constants.as_ptr() as *const u64, n_constants: self.constants.len() as u32,
                rotations: self.rotations.as_ptr(), n_rotations: self.rotations.len() as u32,
                calcs: self.calcs.as_ptr(), n_calcs: self.calcs.len() as u32,
                parts: self.parts.as_ptr(), n_parts: self.parts.len() as u32, n_intermediates: self.n_intermediates }
    }
}

/// A device-resident column (`b200zk_dev_*` handle), freed on drop.
struct DevCol(u64);
impl DevCol {
    fn alloc(n: usize) -> Self { let mut h = 0; ffi::check(unsafe { ffi::b200zk_dev_alloc(n, &mut h) }); DevCol(h) }
    fn from_host(v: &[Fr]) -> Self {
        let c = Self::alloc(v.len());
        ffi::check(unsafe { ffi::b200zk_dev_upload(c.0, 0, v.as_ptr() as *const u64, v.len()) });   // D2D when mirrored
        c
    }
    fn ptr(&self) -> *mut core::ffi::c_void { unsafe { ffi::b200zk_dev_ptr(self.0) } }
}
impl Drop for DevCol { fn drop(&mut self) { unsafe { ffi::b200zk_dev_free(self.0) }; } }

/// What evaluate_h reads from the proving key, uploaded once: `pk.fixed_cosets`, `pk.permutation.cosets`,
/// `pk.l0`, `pk.l_last`, `pk.l_active_row` (all extended-domain columns) and the flattened graphs.
pub(crate) struct DeviceKey {
    fixed: Vec<DevCol>, sigma: Vec<DevCol>, l0: DevCol, l_last: DevCol, l_active_row: DevCol,
    gates: FlatGraph, lookups: Vec<FlatGraph>,
}

impl DeviceKey {
    pub(crate) fn new(pk: &ProvingKey<G1Affine>) -> Self {
        DeviceKey {
            fixed: pk.fixed_cosets.iter().map(|c| DevCol::from_host(c)).collect(),
            sigma: pk.permutation.cosets.iter().map(|c| DevCol::from_host(c)).collect(),
            l0: DevCol::from_host(&pk.l0), l_last: DevCol::from_host(&pk.l_last), l_active_row: DevCol::from_host(&pk.l_active_row),
            gates: FlatGraph::new(&pk.ev.custom_gates), lookups: pk.ev.lookups.iter().map(FlatGraph::new).collect(),
        }
    }
}

impl Evaluator<G1Affine> {
    /// `Evaluator::evaluate_h` — signature unchanged (crate-private upstream as well).
    pub(in crate::plonk) fn evaluate_h(
        &self, pk: &ProvingKey<G1Affine>,
        advice_polys: &[&[Polynomial<Fr, Coeff>]], instance_polys: &[&[Polynomial<Fr, Coeff>]],
        challenges: &[Fr], y: Fr, beta: Fr, gamma: Fr, theta: Fr,
        lookups: &[Vec<lookup::prover::Committed<G1Affine>>], permutations: &[permutation::prover::Committed<G1Affine>],
    ) -> Polynomial<Fr, ExtendedLagrangeCoeff> {
        let domain = &pk.vk.domain;
        let (k, ext_k) = (domain.k(), domain.extended_k());
        let (n, size) = (1usize << k, 1usize << ext_k);
        let dk: &DeviceKey = pk.device_key.get_or_init(|| DeviceKey::new(pk));     // `OnceCell<DeviceKey>` added to ProvingKey
        let (ext_omega, zeta, delta) = (limbs(&domain.get_extended_omega()), limbs(&domain.g_coset), limbs(&Fr::DELTA));
        let handles = |cols: &[DevCol]| cols.iter().map(|c| c.0).collect::<Vec<u64>>();
        // coefficient-form polynomials -> extended columns, one batched device transform per group
        let extend = |polys: &[&Polynomial<Fr, Coeff>]| -> (DevCol, Vec<DevCol>) {
            let stage = DevCol::alloc(polys.len() * n);
            for (i, p) in polys.iter().enumerate() {
                ffi::check(unsafe { ffi::b200zk_dev_upload(stage.0, i * n, p.values.as_ptr() as *const u64, n) });
            }
            let ext = DevCol::alloc(polys.len() * size);
            ffi::check(unsafe {
                ffi::b200zk_coeff_to_extended_dev(stage.ptr(), n, ext.ptr(), size, polys.len(), k, ext_k,
                                                  ext_omega.as_ptr(), zeta.as_ptr(), std::ptr::null_mut())
            });
            let views = (0..polys.len()).map(|i| {
                let mut h = 0;
                ffi::check(unsafe { ffi::b200zk_dev_view(ext.0, i * size, size, &mut h) });
                DevCol(h)
            }).collect();
            (ext, views)
        };
        let values = DevCol::alloc(size);
        let table = DevCol::alloc(size);
        let fixed_h = handles(&dk.fixed);
        let mut first = true;
        for (((advice, instance), lookups), permutation) in
            advice_polys.iter().zip(instance_polys.iter()).zip(lookups.iter()).zip(permutations.iter())
        {
            let (_a_store, advice_ext) = extend(&advice.iter().collect::<Vec<_>>());
            let (_i_store, instance_ext) = extend(&instance.iter().collect::<Vec<_>>());
            let (advice_h, instance_h) = (handles(&advice_ext), handles(&instance_ext));
            let ch: Vec<[u64; 4]> = challenges.iter().map(limbs).collect();
            let env = Env {
                fixed: fixed_h.as_ptr(), n_fixed: fixed_h.len() as u32, advice: advice_h.as_ptr(), n_advice: advice_h.len() as u32,
                instance: instance_h.as_ptr(), n_instance: instance_h.len() as u32,
                challenges: ch.as_ptr() as *const u64, n_challenges: ch.len() as u32,
                beta: limbs(&beta), gamma: limbs(&gamma), theta: limbs(&theta), y: limbs(&y), k, ext_k, range_begin: 0, range_len: 0,
            };
            // custom gates: values[idx] = custom_gates.evaluate(.., &values[idx], ..)
            ffi::check(unsafe { ffi::b200zk_quotient_graph(&dk.gates.as_c(), &env, if first { 0 } else { values.0 }, values.0) });
            first = false;
            // permutation argument (the product cosets were computed by permutation::Argument::commit)
            let sets = &permutation.sets;
            if !sets.is_empty() {
                let p = &pk.vk.cs.permutation;
                let (kind, index): (Vec<u32>, Vec<u32>) = p.columns.iter().map(|c| match c.column_type() {
                    Any::Fixed => (2u32, c.index() as u32), Any::Advice(_) => (3, c.index() as u32), Any::Instance => (4, c.index() as u32),
                }).unzip();
                let products: Vec<DevCol> = sets.iter().map(|s| DevCol::from_host(&s.permutation_product_coset)).collect();
                let (sigma_h, product_h) = (handles(&dk.sigma), handles(&products));
                ffi::check(unsafe {
                    ffi::b200zk_quotient_permutation(&env, values.0, kind.as_ptr(), index.as_ptr(), sigma_h.as_ptr(), kind.len() as u32,
                        product_h.as_ptr(), product_h.len() as u32, (pk.vk.cs_degree - 2) as u32, pk.vk.cs.blinding_factors() as u32,
                        dk.l0.0, dk.l_last.0, dk.l_active_row.0, ext_omega.as_ptr(), zeta.as_ptr(), delta.as_ptr())
                });
            }
            // lookups: three coset transforms and five folded constraints each
            for (n_lookup, lookup) in lookups.iter().enumerate() {
                let (_s, ext) = extend(&[&lookup.product_poly, &lookup.permuted_input_poly, &lookup.permuted_table_poly]);
                ffi::check(unsafe { ffi::b200zk_quotient_graph(&dk.lookups[n_lookup].as_c(), &env, 0, table.0) });
                ffi::check(unsafe {
                    ffi::b200zk_quotient_lookup(&env, values.0, table.0, ext[0].0, ext[1].0, ext[2].0, dk.l0.0, dk.l_last.0, dk.l_active_row.0)
                });
            }
        }
        let mut out = domain.empty_extended();
        ffi::check(unsafe { ffi::b200zk_dev_download(values.0, 0, out.values.as_mut_ptr() as *mut u64, size) });
        out.mark_mirrored();     // b200zk_dev_download left a mirror: divide_by_vanishing_poly / extended_to_coeff find it
        out
    }
}
