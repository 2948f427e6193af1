// Self-test of the CPU verifier's pairing (halo2-lasso_b200/verifier/pairing.hpp), run by tests/test_verifier_cpu.py:
// the split final exponentiation equals the plain (p^12-1)/r power, Fq12 inverse / Frobenius identities, bilinearity.
#include <cstdio>

#include "../../halo2-lasso_b200/verifier/pairing.hpp"
using namespace b200v;

int main() {
  const G1Affine g1 = G1Affine::generator();
  const G2Affine g2 = G2Affine::generator();
  int bad = 0;
  for (uint64_t a = 1; a <= 3; ++a) {
    const G1Affine p = G1::from_affine(g1).mul(Fr::from_u64(1000003 * a + 7)).to_affine();
    const G2Affine q = g2.mul(Fr::from_u64(99991 * a + 5));
    const Fq12 f = miller_loop(p, q);
    bad += !(f * f.inv() == Fq12::one());
    bad += !(final_exponentiation(f) == final_exponentiation_plain(f));
    // x^(p^2) applied six times is x^(p^12) = x
    Fq12 t = f;
    for (int i = 0; i < 6; ++i) t = t.frobenius_p2();
    bad += !(t == f);
    bad += (f.frobenius_p2() == f);
  }
  // bilinearity: e(aP, bQ) == e(abP, Q) == e(P, Q)^(ab), non-degeneracy, and the product check
  const Fr a = Fr::from_u64(123456789), b = Fr::from_u64(987654321);
  const G1Affine ap = G1::from_affine(g1).mul(a).to_affine(), abp = G1::from_affine(g1).mul(a * b).to_affine();
  const G2Affine bq = g2.mul(b);
  bad += !(pairing(ap, bq) == pairing(abp, g2));
  bad += (pairing(g1, g2) == Fq12::one());
  bad += !pairings_product_is_identity({{ap, bq}, {abp, g2.neg()}});
  bad += pairings_product_is_identity({{ap, bq}, {abp, g2}});
  printf("pairing selftest: %d failure(s)\n", bad);
  return bad;
}
