// Host build of the product's limb arithmetic (halo2-lasso_b200/csrc/ff32.cuh) with the emulated
// carry flag, so tests can check the exact PTX-chain algorithm against Python big ints on the CPU.
#include "../../halo2-lasso_b200/csrc/ff32.cuh"
using namespace b200;
extern "C" {
#define API(NAME, P)                                                                          \
  void NAME##_mul(const Fe<P>* a, const Fe<P>* b, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_mul<P>(a[i], b[i]); } \
  void NAME##_add(const Fe<P>* a, const Fe<P>* b, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_add<P>(a[i], b[i]); } \
  void NAME##_sub(const Fe<P>* a, const Fe<P>* b, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_sub<P>(a[i], b[i]); } \
  void NAME##_inv(const Fe<P>* a, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_inv<P>(a[i]); } \
  void NAME##_inv_fermat(const Fe<P>* a, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_inv_fermat<P>(a[i]); } \
  void NAME##_from_canonical(const Fe<P>* a, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_from_canonical<P>(a[i]); } \
  void NAME##_to_canonical(const Fe<P>* a, Fe<P>* o, long n) { for (long i = 0; i < n; ++i) o[i] = fe_to_canonical<P>(a[i]); }
API(h32_fr, FrP)
API(h32_fq, FqP)
}

// Transcript (csrc/transcript.cuh) on the host: write n elements, squeeze, write the challenge,
// write a point, squeeze again. Returns the proof length.
#include "../../halo2-lasso_b200/csrc/transcript.cuh"
extern "C" int h32_transcript_run(const Fr* fes, int n, const Fq* pt_xy, uint8_t* proof, int cap, Fr* ch) {
  Transcript t;
  tr_init(&t, proof, (uint32_t)cap);
  for (int i = 0; i < n; ++i) tr_write_fe(&t, fes[i]);
  ch[0] = tr_squeeze(&t);
  tr_common_fe(&t, ch[0]);
  tr_write_commitment(&t, pt_xy[0], pt_xy[1]);
  ch[1] = tr_squeeze(&t);
  ch[2] = tr_squeeze(&t);
  return t.error ? -1 : (int)t.proof_len;
}

// G1 (csrc/g1.cuh) on the host: Σ k_i * P_i with small k via mul_small / add / add_affine paths.
#include "../../halo2-lasso_b200/csrc/g1.cuh"
extern "C" void h32_g1_lincomb(const G1Aff* pts, const int32_t* ks, int n, G1Aff* out) {
  G1Xyzz acc = g1_identity();
  for (int i = 0; i < n; ++i) {
    int32_t k = ks[i];
    if (k == 1 || k == -1) {
      acc = g1_add_affine(acc, pts[i], k < 0);  // mixed path
    } else {
      G1Xyzz t = g1_mul_small(g1_from_affine(pts[i]), (uint32_t)(k < 0 ? -k : k));
      acc = g1_add(acc, k < 0 ? g1_neg(t) : t);
    }
  }
  *out = g1_to_affine(acc);
}
