// Checked by tests/test_host_cpu.py::test_interleaved_viterbi_table_builder_is_bit_identical: the eight-rows-at-a-time
// builder of the exact Viterbi log-transition tables (titan_host.hpp:viterbi_log_table_rows8, what the library runs)
// against the row-by-row statement of src/fwd_backC_clonalCN.c:283-292,460-465 (viterbi_log_table), bit for bit.
#include "titan_host.hpp"
#include <cstdio>
#include <random>
#include <chrono>
int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(0, 1);
  int bad = 0;
  for (int cfg = 0; cfg < 6; ++cfg) {
    const int U = (cfg % 3 == 0) ? 25 : (cfg % 3 == 1 ? 12 : 36), Z = 1 + cfg, K = U * Z;
    std::vector<unsigned char> het(U, 0);
    if (cfg != 4) het[3] = 1;
    std::vector<unsigned char> cls;
    titan_host::structured_class_matrix(K, U, het.data(), cls);
    std::vector<double> a(4 * K), b(4 * K);
    for (int rep = 0; rep < 300; ++rep) {
      double L = std::pow(10.0, 3 + 12 * u(rng)), d = 1 + std::floor(1e6 * u(rng) * u(rng));
      double rG = titan_host::rho_keep(0, d, L), rZ = titan_host::rho_keep(0, d, L * (rep % 2 ? 1.0 : 5e5));
      if (rep == 0) { rG = 1.0; rZ = 1.0; }
      if (rep == 1) { rG = 0.5; rZ = 0.5; }
      titan_host::viterbi_log_table(K, U, het.data(), rG, rZ, a.data());
      titan_host::viterbi_log_table_rows8(K, cls.data(), rG, rZ, b.data());
      if (std::memcmp(a.data(), b.data(), sizeof(double) * 4 * K) != 0) ++bad;
    }
  }
  // timing at K = 125
  {
    const int U = 25, Z = 5, K = 125;
    std::vector<unsigned char> het(U, 0); het[3] = 1;
    std::vector<unsigned char> cls;
    titan_host::structured_class_matrix(K, U, het.data(), cls);
    std::vector<double> a(4 * K);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < 2000; ++r) titan_host::viterbi_log_table(K, U, het.data(), 1 - 1e-9 * (r + 1), 1 - 1e-4 * (r + 1) / 2000, a.data());
    auto t1 = std::chrono::steady_clock::now();
    for (int r = 0; r < 2000; ++r) titan_host::viterbi_log_table_rows8(K, cls.data(), 1 - 1e-9 * (r + 1), 1 - 1e-4 * (r + 1) / 2000, a.data());
    auto t2 = std::chrono::steady_clock::now();
    std::printf("old %.1f us/pair, new %.1f us/pair\n", std::chrono::duration<double, std::micro>(t1 - t0).count() / 2000,
                std::chrono::duration<double, std::micro>(t2 - t1).count() / 2000);
  }
  std::printf("mismatching tables: %d\n", bad);
  return bad != 0;
}
