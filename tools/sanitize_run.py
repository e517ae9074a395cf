"""Small end-to-end run for compute-sanitizer: every kernel of the library on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import __graft_entry__ as g
g.build()
import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from muscle_synergies_b200.pipeline import trial_synergies
from muscle_synergies_b200 import analysis, emg
from tools.synth_vicon import synth_layout

golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
data = ms.load_vicon_file(os.path.join(golden, "abridged_data.csv"))
print("abridged", data.emg.df.shape)
for name in sorted(os.listdir(os.path.join(golden, "variants")))[:40]:
    try:
        ms.load_vicon_file(os.path.join(golden, "variants", name))
    except Exception as exc:  # noqa: BLE001 - the variants include malformed files
        pass
blob = synth_layout("D", seed=0)
loader = ms.ViconLoader()
d = torch.empty(loader.padded_size(blob.nbytes), dtype=torch.uint8, device="cuda")
d[: blob.nbytes].copy_(torch.from_numpy(blob))
trial = loader.load_device(d, n=blob.nbytes, name="D", defer_check=True)
seg = Segmenter(trial, cut_phases_of=(trial.emg, trial.traj[0]))
print("transitions", len(seg.transitions), "cuts", len(seg.phase_cuts(trial.emg)))
res = trial_synergies(trial, 1, 5, n_restarts=2, max_iter=20, segmenter=seg)
print("pipeline", len(res.cycles))
x = np.abs(np.random.default_rng(0).normal(size=(3000, 6)))
analysis.nmf_mu_batched(x, [2, 3, 4], [0, 1, 2], max_iter=10, regime="stream")
df = pd.DataFrame(x, columns=list("abcdef"))
emg.linear_envelope(df, 6.0, 2000, 4); emg.rms(df, 100); emg.digital_filter(df, (20.0, 450.0), 2000, 2, band_type="bandpass", zero_lag=False)
torch.cuda.synchronize()
print("sanitize run done")
