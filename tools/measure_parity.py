"""Print the MEASURED fp32 parity errors behind the tolerances in tests/test_rollout_gpu.py (run on the GPU box):

    python tools/measure_parity.py > gpurun_out/parity.json

For every forward golden: rel. error of each output vs the reference's fp32 CPU result.  For the gradient envelope
cases: error of the fp32 adjoint and of the reference's own fp32 autograd (the oracle) against fp64 autograd.
The tests hold each quantity to <= 3x (goldens) / <= 10x (adjoint) of what is printed here.
"""
import json
import os
import sys

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [_R, os.path.join(_R, "tests")]
import torch  # noqa: E402
from helpers_mfb import load_golden, rel_err  # noqa: E402
import test_rollout_gpu as T  # noqa: E402


def goldens():
    out = {}
    for name in list(T.FWD_GOLDENS) + ["marv_hill128_joints_T60_B2"]:
        g = load_golden(name)
        if "joint_angles" in g:
            sim, cfg = T._module("marv", float(g["grid_res"]), int(g["T"]), "step")
            B = g["controls"].shape[0]
            with torch.no_grad():
                (states, forces) = sim(T._t(g["z"]).unsqueeze(0).expand(B, -1, -1), T._t(g["controls"]),
                                       joint_angles=T._t(g["joint_angles"]))
        else:
            (states, forces), cfg = T._run_golden(g)
        Xs, Xds, Rs, Oms = states
        Fs, Ff = forces
        keep = g["F_keep_steps"]
        out[name] = {"Xs": rel_err(Xs, g["Xs"]), "Rs": rel_err(Rs, g["Rs"]), "Xds": rel_err(Xds, g["Xds"]),
                     "Omegas": rel_err(Oms, g["Omegas"]), "Fs_keep": rel_err(Fs[:, keep], g["Fs_keep"]),
                     "Ff_keep": rel_err(Ff[:, keep], g["Ff_keep"]), "Fs_sum": rel_err(Fs.double().sum(dim=2), g["Fs_sum"])}
    return out


def main():
    res = {"goldens": goldens()}
    res["grad_envelope"] = {c: T.gradient_envelope(c) for c in T.GRAD_ENVELOPE_CASES}
    res["adjoint_T50"] = {k: T.fp32_adjoint_errors(tape) for k, tape in (("sweep", True), ("three_pass", False))}
    res["bench_inputs_forward"] = {c: T.bench_forward_envelope(c) for c in ("hill", "flat")}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
