/* A host program in plain C99 that uses nothing but the C ABI (include/b200_lasso.h): what a cgo / Rust-FFI caller
 * links against. It runs the two headline workloads on synthetic inputs drawn from the documented splitmix64
 * stream and writes the proof bytes to files, so that tests/test_gpu_cabi.py can compare them with the oracle:
 *
 *   cabi_demo <out_dir> [mu]
 *     sumcheck.bin  cfg2 shape: ClassicSumCheck of eq(x,y) * a(x) * b(x), n = 12, claimed sum "one"
 *     lasso.bin     cfg3 shape: 64-bit range check via Surge, c = 4 chunks of 16 bits, 2^mu lookups (default 8)
 *
 * Build: gcc -std=c99 -O2 -Iinclude examples/cabi_demo.c -Lhalo2-lasso_b200 -lb200lasso -Wl,-rpath,$PWD/halo2-lasso_b200
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200_lasso.h"

static uint64_t sm64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
/* n canonical 253-bit integers: element i = limbs sm64(seed, 4i .. 4i+3), top limb masked to 61 bits */
static void rand_canonical(uint64_t seed, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < 4 * n; ++i) out[i] = sm64(seed, i);
  for (uint64_t i = 0; i < n; ++i) out[4 * i + 3] &= 0x1FFFFFFFFFFFFFFFULL;
}
#define CHECK(call)                                                      \
  do {                                                                   \
    int rc_ = (call);                                                    \
    if (rc_ != B200_OK) {                                                \
      fprintf(stderr, "%s failed with status %d\n", #call, rc_);         \
      return 1;                                                          \
    }                                                                    \
  } while (0)

/* canonical integers -> Montgomery residues, converted on the device (the library owns the field arithmetic) */
static int to_montgomery(b200_ctx* ctx, uint64_t* vals, uint64_t n) {
  void* dev = NULL;
  CHECK(b200_poly_upload(ctx, vals, n, &dev));
  CHECK(b200_fr_convert(ctx, dev, dev, n, 1));
  CHECK(b200_poly_download(ctx, dev, n, vals));
  CHECK(b200_poly_free(ctx, dev));
  return 0;
}
static int write_proof(b200_ctx* ctx, const char* dir, const char* name) {
  static uint8_t buf[1 << 20];
  uint64_t len = 0;
  char path[1024];
  CHECK(b200_transcript_proof(ctx, buf, sizeof buf, &len));
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  fwrite(buf, 1, len, f);
  fclose(f);
  printf("%s: %llu proof bytes\n", name, (unsigned long long)len);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <out_dir> [mu]\n", argv[0]);
    return 2;
  }
  const char* dir = argv[1];
  const int mu = argc > 2 ? atoi(argv[2]) : 8;
  b200_ctx* ctx = NULL;
  CHECK(b200_ctx_create(0, &ctx));

  /* ---- cfg2 shape: sum-check of eq * a * b ---------------------------------------------------------- */
  {
    enum { N_VARS = 12 };
    const uint64_t n = 1ull << N_VARS;
    uint64_t* a = malloc(32 * n);
    uint64_t* b = malloc(32 * n);
    uint64_t y[4 * N_VARS], one[4] = {1, 0, 0, 0}, challenges[4 * N_VARS], evals[8];
    rand_canonical(1, n, a);
    rand_canonical(2, n, b);
    rand_canonical(3, N_VARS, y);
    if (to_montgomery(ctx, a, n) || to_montgomery(ctx, b, n) || to_montgomery(ctx, y, N_VARS) || to_montgomery(ctx, one, 1))
      return 1;
    const void* tables[2] = {a, b};
    CHECK(b200_transcript_reset(ctx));
    CHECK(b200_sumcheck_prove_evals_host(ctx, N_VARS, 1, 2, tables, one, y, one, challenges, evals));
    if (write_proof(ctx, dir, "sumcheck.bin")) return 1;
    free(a);
    free(b);
  }

  /* ---- cfg3 shape: Lasso range check ---------------------------------------------------------------- */
  {
    const int srs_vars = mu > 16 ? mu : 16;
    uint64_t ss[4 * 32];
    rand_canonical(7, (uint64_t)srs_vars, ss);
    if (to_montgomery(ctx, ss, (uint64_t)srs_vars)) return 1;
    CHECK(b200_kzg_setup(ctx, ss, srs_vars));
    const uint64_t m = 1ull << mu;
    uint64_t* xs = malloc(8 * m);
    for (uint64_t i = 0; i < m; ++i) xs[i] = sm64(5, i);
    for (uint64_t i = m / 2; i < m; ++i) xs[i] = xs[i - m / 2]; /* repeated addresses (read_ts != 0) */
    CHECK(b200_transcript_reset(ctx));
    CHECK(b200_lasso_prove(ctx, 0 /* range */, 4, mu, xs, NULL));
    if (write_proof(ctx, dir, "lasso.bin")) return 1;
    free(xs);
  }
  CHECK(b200_sync(ctx));
  printf("kernel launches: %llu\n", (unsigned long long)b200_launch_count(ctx, 0));
  b200_ctx_destroy(ctx);
  return 0;
}
