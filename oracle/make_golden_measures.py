"""TEST INFRASTRUCTURE ONLY — pin the fused reward functors against the UNMODIFIED reference's ``QCMSolver``.

Run in the build container (where /root/reference exists):  ``python oracle/make_golden_measures.py``.
Input: the reference trajectory ``E5_States_dop853_tight`` of ``tests/golden/lyap_optomech.npz`` (itself produced by the
reference, see make_golden.py) plus a noisy copy (so that asymmetric covariance blocks are exercised as well).
Output ``tests/golden/measures_optomech.npz``: for both inputs the reference's own
``get_entanglement_logarithmic_negativity_2`` (quantrl/solvers/measure.py:401-440), ``get_synchronization_complete``
(:442-468) and ``get_synchronization_phase`` (:470-514) with ``pos_i = 0, pos_j = 2``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
ref_shim.load()
from quantrl.solvers.measure import QCMSolver  # noqa: E402


def measures(St, m=2, q=4):
    Corrs = St[..., 2 * m:].reshape(-1, q, q)
    Modes = (St[..., :m] + 1j * St[..., m:2 * m]).reshape(-1, m)
    qcm = QCMSolver(Modes=Modes, Corrs=Corrs, params={"show_progress": False, "measure_codes": ["entan_ln_2"], "indices": (0, 1)})
    shape = St.shape[:-1]
    return {"entan_ln": qcm.get_entanglement_logarithmic_negativity_2(pos_i=0, pos_j=2).reshape(shape),
            "sync_c": qcm.get_synchronization_complete(pos_i=0, pos_j=2).reshape(shape),
            "sync_p": qcm.get_synchronization_phase(pos_i=0, pos_j=2).reshape(shape)}


if __name__ == "__main__":
    g = np.load(os.path.join(GOLDEN, "lyap_optomech.npz"))
    St = g["E5_States_dop853_tight"]
    noisy = St + 1e-2 * np.random.default_rng(7).standard_normal(St.shape)
    out = {"States": St, "States_noisy": noisy}
    out.update({k + "_ref": v for k, v in measures(St).items()})
    out.update({k + "_noisy_ref": v for k, v in measures(noisy).items()})
    np.savez_compressed(os.path.join(GOLDEN, "measures_optomech.npz"), **out)
    print({k: (np.shape(v), float(np.min(v)), float(np.max(v))) for k, v in out.items()})
