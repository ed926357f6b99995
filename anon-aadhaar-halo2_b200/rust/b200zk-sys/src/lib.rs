//! Raw FFI declarations of include/b200zk.h (hand-written; keep in sync with the header —
//! tests/test_abi.py checks the header against the shared library's exports).
//!
//! NOT COMPILED HERE: no Rust toolchain exists in this image.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct b200zk_src { pub kind: u32, pub a: u32, pub b: u32 }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct b200zk_calc { pub op: u32, pub target: u32, pub x: b200zk_src, pub y: b200zk_src, pub parts_off: u32, pub parts_len: u32 }

#[repr(C)]
pub struct b200zk_graph {
    pub constants: *const u64, pub n_constants: u32,
    pub rotations: *const i32, pub n_rotations: u32,
    pub calcs: *const b200zk_calc, pub n_calcs: u32,
    pub parts: *const b200zk_src, pub n_parts: u32,
    pub n_intermediates: u32,
}

#[repr(C)]
pub struct b200zk_quotient_env {
    pub fixed: *const u64, pub n_fixed: u32,
    pub advice: *const u64, pub n_advice: u32,
    pub instance: *const u64, pub n_instance: u32,
    pub challenges: *const u64, pub n_challenges: u32,
    pub beta: [u64; 4], pub gamma: [u64; 4], pub theta: [u64; 4], pub y: [u64; 4],
    pub k: u32, pub ext_k: u32,
    pub range_begin: u64, pub range_len: u64,
}

extern "C" {
    pub fn b200zk_init(device: c_int) -> c_int;
    pub fn b200zk_shutdown() -> c_int;
    pub fn b200zk_last_error() -> *const c_char;
    pub fn b200zk_abi_version() -> u32;
    pub fn b200zk_stream_release(stream: *mut c_void) -> c_int;
    pub fn b200zk_mirror_enable(max_bytes: usize) -> c_int;
    pub fn b200zk_mirror_invalidate(host_ptr: *const c_void, bytes: usize) -> c_int;
    pub fn b200zk_mirror_stats(out: *mut u64) -> c_int;

    pub fn b200zk_ntt(a: *mut u64, log_n: u32, omega: *const u64) -> c_int;
    pub fn b200zk_intt(a: *mut u64, log_n: u32, omega_inv: *const u64, divisor: *const u64) -> c_int;
    pub fn b200zk_coeff_to_extended(input: *const u64, k: u32, out: *mut u64, ext_k: u32, extended_omega: *const u64, zeta: *const u64) -> c_int;
    pub fn b200zk_extended_to_coeff(a: *const u64, ext_k: u32, extended_omega_inv: *const u64, extended_ifft_divisor: *const u64, zeta: *const u64, out: *mut u64, keep: usize) -> c_int;
    pub fn b200zk_divide_by_vanishing(h: *mut u64, ext_k: u32, t_evaluations: *const u64, t_len: u32) -> c_int;
    pub fn b200zk_ntt_many(a: *mut u64, stride: usize, count: usize, log_n: u32, omega: *const u64) -> c_int;
    pub fn b200zk_intt_many(a: *mut u64, stride: usize, count: usize, log_n: u32, omega_inv: *const u64, divisor: *const u64) -> c_int;
    pub fn b200zk_coeff_to_extended_many(input: *const u64, in_stride: usize, out: *mut u64, out_stride: usize, count: usize, k: u32, ext_k: u32, extended_omega: *const u64, zeta: *const u64) -> c_int;
    pub fn b200zk_ntt_dev(d_a: *mut c_void, stride: usize, count: usize, log_n: u32, omega: *const u64, divisor_or_null: *const u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_coeff_to_extended_dev(d_in: *const c_void, in_stride: usize, d_out: *mut c_void, out_stride: usize, count: usize, k: u32, ext_k: u32, extended_omega: *const u64, zeta: *const u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_extended_to_coeff_dev(d_a: *const c_void, ext_k: u32, extended_omega_inv: *const u64, extended_ifft_divisor: *const u64, zeta: *const u64, d_t_evaluations_or_null: *const c_void, t_len: u32, d_out: *mut c_void, keep: usize, stream: *mut c_void) -> c_int;

    pub fn b200zk_coeff_to_coset_dev(d_in: *const c_void, in_stride: usize, d_out: *mut c_void, out_stride: usize, count: usize, k: u32, omega: *const u64, coset_generator: *const u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_extended_coset_slice_dev(d_ext: *const c_void, ext_stride: usize, d_out: *mut c_void, out_stride: usize, count: usize, k: u32, ext_k: u32, coset: u32, stream: *mut c_void) -> c_int;
    pub fn b200zk_extended_coset_interleave_dev(d_coset: *const c_void, d_ext: *mut c_void, k: u32, ext_k: u32, coset: u32, stream: *mut c_void) -> c_int;
    pub fn b200zk_ntt4_first_pass_scatter_dev(d_in: *const c_void, log_n: u32, log_n1: u32, omega: *const u64, world: u32, rank: u32, dest_bases: *const *mut c_void, dest_pitch: usize, dest_col_offset: usize, stream: *mut c_void) -> c_int;
    pub fn b200zk_ntt4_twiddle_scatter_dev(d_in: *const c_void, log_n: u32, log_n1: u32, omega: *const u64, world: u32, rank: u32, dest_bases: *const *mut c_void, dest_pitch: usize, dest_col_offset: usize, stream: *mut c_void) -> c_int;
    pub fn b200zk_ntt4_gather_rows_dev(d_recv: *const c_void, d_out: *mut c_void, log_n: u32, log_n1: u32, world: u32, stream: *mut c_void) -> c_int;
    pub fn b200zk_msm_g1(scalars: *const u64, bases: *const u64, n: usize, out_xyz: *mut u64) -> c_int;
    pub fn b200zk_bases_register(bases: *const u64, n: usize, handle_out: *mut u64) -> c_int;
    pub fn b200zk_bases_register_ex(bases: *const u64, n: usize, precompute_windows: c_int, handle_out: *mut u64) -> c_int;
    pub fn b200zk_bases_evict(handle: u64) -> c_int;
    pub fn b200zk_msm_g1_registered_many(handle: u64, scalars: *const u64, stride: usize, count: usize, n: usize, out_xyz: *mut u64) -> c_int;
    pub fn b200zk_msm_g1_registered_dev(handle: u64, d_scalars: *const c_void, stride: usize, count: usize, n: usize, d_out_xyz: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b200zk_msm_g1_registered(handle: u64, scalars: *const u64, n: usize, out_xyz: *mut u64) -> c_int;
    pub fn b200zk_msm_g1_dev(d_scalars: *const c_void, d_bases: *const c_void, n: usize, out_xyz: *mut u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_msm_g1_dev_async(d_scalars: *const c_void, d_bases: *const c_void, n: usize, d_out_xyz: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b200zk_g1_sum(points_xyz: *const u64, count: usize, out_xyz: *mut u64) -> c_int;

    pub fn b200zk_batch_invert(a: *mut u64, n: usize) -> c_int;
    pub fn b200zk_batch_invert_dev(d_a: *mut c_void, n: usize, stream: *mut c_void) -> c_int;
    pub fn b200zk_prefix_product_dev(d_in: *const c_void, in_stride: usize, d_out: *mut c_void, out_stride: usize, count: usize, n: usize, first_or_null: *const u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_permutation_product_dev(d_values: *const *const c_void, d_sigma: *const *const c_void, n_cols: u32, chunk_len: u32, k: u32, beta: *const u64, gamma: *const u64, omega: *const u64, delta: *const u64, blinding_factors: u32, blinds_or_null: *const u64, d_z: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b200zk_lookup_product_dev(d_compressed_input: *const *const c_void, d_compressed_table: *const *const c_void, d_permuted_input: *const *const c_void, d_permuted_table: *const *const c_void, count: u32, k: u32, beta: *const u64, gamma: *const u64, blinding_factors: u32, blinds_or_null: *const u64, d_z: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b200zk_permute_expression_pair_dev(d_input: *const c_void, d_table: *const c_void, stride: usize, count: u32, k: u32, blinding_factors: u32, blinds_or_null: *const u64, d_permuted_input: *mut c_void, d_permuted_table: *mut c_void, out_stride: usize, stream: *mut c_void) -> c_int;
    pub fn b200zk_linear_combination_dev(d_polys: *const *const c_void, coeffs: *const u64, count: u32, n: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b200zk_eval_polynomial_dev(d_polys: *const c_void, stride: usize, count: usize, n: usize, points: *const u64, out: *mut u64, stream: *mut c_void) -> c_int;
    pub fn b200zk_kate_division_dev(d_a: *const c_void, n: usize, b: *const u64, d_q: *mut c_void, stream: *mut c_void) -> c_int;

    pub fn b200zk_g1_to_bytes(points_xyz: *const u64, count: usize, out32: *mut u8) -> c_int;
    pub fn b200zk_g1_affine_to_bytes(points_xy: *const u64, count: usize, out32: *mut u8) -> c_int;
    pub fn b200zk_g1_to_evm_bytes(points_xyz: *const u64, count: usize, out64: *mut u8) -> c_int;
    pub fn b200zk_g1_affine_from_bytes(in32: *const u8, count: usize, points_xy: *mut u64) -> c_int;

    // measurement and test support
    pub fn b200zk_gen_scalars_dev(d_out: *mut c_void, n: usize, seed: u64, start: usize) -> c_int;
    pub fn b200zk_gen_points_dev(d_out: *mut c_void, n: usize, seed: u64, start: usize) -> c_int;
    pub fn b200zk_field_op(field: u32, op: u32, a: *const u64, b: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn b200zk_modmul_peak(iters: u32, modmul_per_s_out: *mut f64) -> c_int;
    pub fn b200zk_msm_profile(enable: c_int) -> c_int;
    pub fn b200zk_msm_tune(max_chunk: u32, max_seglen: u32, force_window_bits: u32) -> c_int;
    pub fn b200zk_msm_last_stages(ms_out: *mut f32, capacity: c_int, info_out: *mut u64) -> c_int;
    pub fn b200zk_msm_upload_pipeline(parts: u32, min_n: usize) -> c_int;
    pub fn b200zk_msm_upload_ranges(n: usize, parts: u32, growth: f64, begin_out: *mut usize, count_out: *mut u32) -> c_int;
    pub fn b200zk_ntt_transfer_pipeline(chunks: u32, min_log_n: u32) -> c_int;
    pub fn b200zk_ntt_tune(direct_twiddle_max_log_n: u32) -> c_int;
    pub fn b200zk_kernel_launches() -> u64;

    // page-locked host memory
    pub fn b200zk_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn b200zk_host_unregister(ptr: *mut c_void) -> c_int;
    pub fn b200zk_host_alloc(bytes: usize, ptr_out: *mut *mut c_void) -> c_int;
    pub fn b200zk_host_free(ptr: *mut c_void) -> c_int;

    pub fn b200zk_dev_alloc(n_elems: usize, handle_out: *mut u64) -> c_int;
    pub fn b200zk_dev_free(handle: u64) -> c_int;
    pub fn b200zk_dev_view(parent: u64, offset: usize, n_elems: usize, handle_out: *mut u64) -> c_int;
    pub fn b200zk_dev_upload(handle: u64, offset: usize, host: *const u64, n_elems: usize) -> c_int;
    pub fn b200zk_dev_download(handle: u64, offset: usize, host: *mut u64, n_elems: usize) -> c_int;
    pub fn b200zk_dev_ptr(handle: u64) -> *mut c_void;

    pub fn b200zk_quotient_graph(graph: *const b200zk_graph, env: *const b200zk_quotient_env, previous_handle: u64, out_handle: u64) -> c_int;
    pub fn b200zk_quotient_permutation(env: *const b200zk_quotient_env, values_handle: u64, column_kind: *const u32, column_index: *const u32, sigma_handles: *const u64, n_columns: u32, product_handles: *const u64, n_sets: u32, chunk_len: u32, blinding_factors: u32, l0_handle: u64, l_last_handle: u64, l_active_row_handle: u64, extended_omega: *const u64, zeta: *const u64, delta: *const u64) -> c_int;
    pub fn b200zk_quotient_lookup(env: *const b200zk_quotient_env, values_handle: u64, table_values_handle: u64, product_handle: u64, permuted_input_handle: u64, permuted_table_handle: u64, l0_handle: u64, l_last_handle: u64, l_active_row_handle: u64) -> c_int;
}

/// Panic with the library's message, as upstream's infallible functions would on a failed assert.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(b200zk_last_error()) }.to_string_lossy().into_owned();
        panic!("b200zk: {msg}");
    }
}
