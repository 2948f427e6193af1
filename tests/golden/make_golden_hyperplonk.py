"""Generates tests/golden/hyperplonk_golden.json with the pure-Python HyperPlonk model (pymodel_hyperplonk.py):
whole proofs of the reference's two test circuits (vanilla plonk with one and two permutation chunks, vanilla plonk
with the LogUp lookup) at k = 3 and of the two-instance-column / two-phase circuit at k = 4, as committed bytes.
Inputs are the seeded fixtures of halo2-lasso_b200/hyperplonk.py; about ten seconds of CPU."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import pymodel as M
import pymodel_hyperplonk as MH
from halo2_lasso_b200 import hyperplonk as H

cases = []
for lookup, max_degree in ((False, 4), (True, 4), (False, 3)):
    k, seed = 3, 31
    info, instances, w = (H.rand_vanilla_plonk_with_lookup_circuit if lookup else H.rand_vanilla_plonk_circuit)(k, seed, num_instances=2)
    proof = MH.prove(M.kzg_setup(M.rand_fr(7, k)), info, instances, w, max_degree=max_degree)
    cases.append({"circuit": "vanilla_plonk_with_lookup" if lookup else "vanilla_plonk", "k": k, "seed": seed, "srs_seed": 7,
                  "max_degree": max_degree, "proof": proof.hex()})
for with_lookup in (True, False):
    k, seed = 4, 61
    info, inst_cols, synth = H.rand_two_phase_circuit(k, seed, with_lookup)
    proof = MH.prove(M.kzg_setup(M.rand_fr(7, k)), info, inst_cols, synth)
    cases.append({"circuit": "two_phase", "k": k, "seed": seed, "srs_seed": 7, "with_lookup": with_lookup, "proof": proof.hex()})
json.dump({"cases": cases}, open(os.path.join(HERE, "hyperplonk_golden.json"), "w"), indent=1)
print("written", [(c["circuit"], len(c["proof"]) // 2) for c in cases])
