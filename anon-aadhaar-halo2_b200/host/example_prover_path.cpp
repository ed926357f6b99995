// Drives the hot path the way create_proof does for one advice column, through the C++
// mirror: commit_lagrange -> lagrange_to_coeff -> coeff_to_extended -> (vanishing divide)
// -> extended_to_coeff -> commit.  Inputs come from a file of raw limbs; results are
// printed as hex limbs so tests/test_host_cpp.py can compare them with the oracle.
//   usage: example_prover_path <k> <j> <scalars.bin> <bases.bin>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "halo2_b200.hpp"

template <class T> static std::vector<T> read_vec(const char* path, size_t count) {
    std::vector<T> v(count);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(T)));
    if (!f) throw std::runtime_error(std::string("short read: ") + path);
    return v;
}
template <class T> static void dump(const char* tag, const T& limbs) {
    std::printf("%s", tag);
    for (uint64_t l : limbs) std::printf(" %016llx", (unsigned long long)l);
    std::printf("\n");
}
static void dump_vec_digest(const char* tag, const std::vector<halo2::Fr>& v) {
    // order-sensitive 64-bit digest (FNV-1a over the limbs) + first / last element
    uint64_t h = 1469598103934665603ull;
    for (const auto& e : v) for (uint64_t l : e) { h ^= l; h *= 1099511628211ull; }
    std::printf("%s %zu %016llx\n", tag, v.size(), (unsigned long long)h);
    dump("  first", v.front());
    dump("  last", v.back());
}

int main(int argc, char** argv) {
    try {
        if (argc == 4 && std::string(argv[1]) == "--constants") {
            // host-only: EvaluationDomain::new constants (no device needed)
            halo2::EvaluationDomain d((uint32_t)std::atoi(argv[3]), (uint32_t)std::atoi(argv[2]));
            dump("omega", d.omega); dump("omega_inv", d.omega_inv);
            dump("extended_omega", d.extended_omega); dump("extended_omega_inv", d.extended_omega_inv);
            dump("g_coset", d.g_coset); dump("g_coset_inv", d.g_coset_inv);
            dump("ifft_divisor", d.ifft_divisor); dump("extended_ifft_divisor", d.extended_ifft_divisor);
            std::printf("extended_k %u t_len %zu\n", d.extended_k, d.t_evaluations.size());
            for (const auto& t : d.t_evaluations) dump("t", t);
            return 0;
        }
        if (argc != 5) { std::fprintf(stderr, "usage: %s k j scalars.bin bases.bin\n", argv[0]); return 2; }
        const uint32_t k = (uint32_t)std::atoi(argv[1]), j = (uint32_t)std::atoi(argv[2]);
        const size_t n = (size_t)1 << k;
        auto column = read_vec<halo2::Fr>(argv[3], n);
        auto g_lagrange = read_vec<halo2::G1Affine>(argv[4], n);
        halo2::check(b200zk_init(0));
        halo2::EvaluationDomain domain(j, k);
        dump("omega", domain.omega);
        dump("extended_omega", domain.extended_omega);
        std::printf("extended_k %u t_len %zu\n", domain.extended_k, domain.t_evaluations.size());
        halo2::ParamsKZG params(g_lagrange, g_lagrange);
        halo2::PageLocked<halo2::Fr> pin_column(column);   // the witness column lives for the whole proof
        dump("commit_lagrange", params.commit_lagrange(column));
        auto coeff = domain.lagrange_to_coeff(column);
        dump_vec_digest("lagrange_to_coeff", coeff);
        auto ext = domain.coeff_to_extended(coeff);
        dump_vec_digest("coeff_to_extended", ext);
        auto hdiv = domain.divide_by_vanishing_poly(ext);
        dump_vec_digest("divide_by_vanishing_poly", hdiv);
        auto back = domain.extended_to_coeff(ext);
        dump_vec_digest("extended_to_coeff", back);
        dump("commit", params.commit(coeff));
        dump("best_multiexp", halo2::best_multiexp(coeff, g_lagrange));
        auto fwd = coeff;
        halo2::best_fft(fwd, domain.omega, k);
        std::printf("best_fft_roundtrip %d\n", fwd == column ? 1 : 0);
        // error behaviour: wrong length is an assertion failure, as upstream
        try {
            std::vector<halo2::Fr> bad(n - 1);
            halo2::best_fft(bad, domain.omega, k);
            std::printf("length_assert 0\n");
        } catch (const std::runtime_error&) { std::printf("length_assert 1\n"); }
        b200zk_shutdown();
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
