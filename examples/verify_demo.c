/* Plain C99 host program over include/b200_verify.h (libb200verify.so: CPU only, no CUDA, no Python): verifies a Lasso
 * proof file as written by examples/cabi_demo.c or by `Keccak256Transcript::into_proof`.
 *
 *   verify_demo <proof file> <trapdoor file: num_vars x 32-byte Montgomery scalars> <kind> <chunks> <mu>
 *
 * exit status 0 = accepted, 1 = rejected, 2 = usage / argument error. Build:
 *   gcc -std=c99 -O2 -Iinclude examples/verify_demo.c -Lhalo2-lasso_b200 -lb200verify -Wl,-rpath,$PWD/halo2-lasso_b200 */
#include <stdio.h>
#include <stdlib.h>

#include "b200_verify.h"

static unsigned char* slurp(const char* path, long* len) {
  FILE* f = fopen(path, "rb");
  unsigned char* buf;
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  *len = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf = (unsigned char*)malloc(*len > 0 ? (size_t)*len : 1);
  if (buf && fread(buf, 1, (size_t)*len, f) != (size_t)*len) {
    free(buf);
    buf = NULL;
  }
  fclose(f);
  return buf;
}

int main(int argc, char** argv) {
  long proof_len = 0, ss_len = 0;
  unsigned char *proof, *ss;
  b200v_kzg* vp = NULL;
  b200v_transcript* tr = NULL;
  int rc;
  if (argc != 6) {
    fprintf(stderr, "usage: %s proof.bin trapdoor.bin kind chunks mu\n", argv[0]);
    return 2;
  }
  proof = slurp(argv[1], &proof_len);
  ss = slurp(argv[2], &ss_len);
  if (!proof || !ss || ss_len % 32 != 0) return 2;
  if (b200v_kzg_setup(ss, (int)(ss_len / 32), &vp) != B200V_ACCEPT) return 2;
  if (b200v_transcript_new(proof, (uint64_t)proof_len, &tr) != B200V_ACCEPT) return 2;
  rc = b200v_lasso_verify(vp, tr, atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  if (rc == B200V_ACCEPT) rc = b200v_transcript_done(tr); /* nothing may follow the Lasso section here */
  printf("%s (%ld proof bytes)\n", rc == B200V_ACCEPT ? "accepted" : rc == B200V_REJECT ? "rejected" : "argument error", proof_len);
  b200v_transcript_free(tr);
  b200v_kzg_free(vp);
  free(proof);
  free(ss);
  return rc;
}
