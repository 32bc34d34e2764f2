"""Regenerates tests/golden/ from the reference tree (run in the build container only; the GPU
box has no /root/reference). Golden vectors the reference ships for this path
(SURVEY 8c): Examples/PremadeMatrix/{Hamiltonian,Overlap}.mtx -> Density-Reference.mtx.
The .mtx files are DATA fixtures (known-answer vectors), copied verbatim.
Also stores the oracle's solver trace for that case (iteration counts, energy, chemical
potential) so that `-m "not gpu"` pins the oracle and `-m gpu` pins the CUDA path to it."""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Examples/PremadeMatrix"


def main():
    for f in ("Hamiltonian.mtx", "Overlap.mtx", "Density-Reference.mtx"):
        shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, "premade_" + f))
    shutil.copyfile("/root/reference/Examples/ComplexMatrix/input.mtx", os.path.join(HERE, "complex_input.mtx"))
    import numpy as np
    import scipy.io as sio
    from oracle import oracle as O
    H = O.PSMatrix.from_scipy(sio.mmread(os.path.join(HERE, "premade_Hamiltonian.mtx")))
    S = O.PSMatrix.from_scipy(sio.mmread(os.path.join(HERE, "premade_Overlap.mtx")))
    D = sio.mmread(os.path.join(HERE, "premade_Density-Reference.mtx")).toarray()
    p = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQ, i1 = O.inverse_square_root(S, p)
    p.converge_diff = 1e-5
    out = {"isq_loop_counter": i1.iterations}
    for name, fn in (("trs2", O.trs2), ("trs4", O.trs4), ("pm", O.pm)):
        K, info = fn(H, ISQ, 5.0, p)
        out[name] = {"loop_counter": info.iterations, "energy": info.energy, "mu": info.chemical_potential,
                     "err_vs_reference_density": float(np.linalg.norm(K.todense() - D))}
    json.dump(out, open(os.path.join(HERE, "premade_oracle_trace.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
