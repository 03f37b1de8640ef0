#!/bin/bash
# The reference's own train.py / test.py on the B200-native path, end to end, with a log for profiles/:
#   python tools/install_reference.py && gpurun -- bash tools/dropin_demo.sh
# 120 iterations crossing both phase boundaries of the schedule (radiance-only < 40 <= full IBL < 80 <= priors + freeze),
# checkpoint + test-set export at iteration 100, then test.py on the checkpoint.
set -e
cd "$(dirname "$0")/.."
T=${TMPDIR:-/tmp}/ibln_dropin_demo
rm -rf "$T"; mkdir -p "$T" gpurun_out
python -m ibl_nerf_b200.synthetic_dataset "$T/data/kitchen" --views 6 --test-views 2 --size 96 128 > "$T/dataset.json"
S=$(date +%s.%N)
python -m ibl_nerf_b200.launcher baseline/_ref/src train --config ../configs/IBL-NeRF/kitchen/IBL-NeRF.txt --datadir "$T/data/kitchen" \
  --basedir "$T/logs" --N_iter 120 --N_iter_ignore_approximated_radiance 40 --N_iter_ignore_prior 80 --N_rand 4096 --chunk 32768 \
  --i_testset 100 --i_weights 100 --summary_step 10 > "$T/train.out" 2> "$T/train.err"
E=$(date +%s.%N)
python -m ibl_nerf_b200.launcher baseline/_ref/src test --config ../configs/IBL-NeRF/kitchen/IBL-NeRF.txt --datadir "$T/data/kitchen" \
  --basedir "$T/logs" --chunk 32768 > "$T/test.out" 2> "$T/test.err"
python - "$T" "$S" "$E" <<'PY' | tee gpurun_out/r2_dropin_demo.txt
import glob, os, sys
from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
T, S, E = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
exp = os.path.join(T, "logs", "IBL-NeRF")
acc = EventAccumulator(exp, size_guidance={"scalars": 0}); acc.Reload()
print("reference src/train.py + src/test.py (unmodified, baseline/_ref) through ibl_nerf_b200.launcher; synthetic Mitsuba-format scene 96x128, 6 views")
print("train.py: 120 iterations, N_rand 4096, phases: radiance-only < 40 <= full IBL < 80 <= priors + freeze; wall %.1f s incl. dataset load, test-set export at 100" % (E - S))
el = [(e.step, e.value) for e in acc.Scalars("elapsed_time")]
if len(el) >= 3:      # train.py accumulates the wall time of its iterations (train.py:501-502); skip the first interval (warm-up)
    print("train.py iteration time through the drop-in (autograd route, the driver's own torch losses / optimizer): "
          + ", ".join("%d-%d: %.1f ms" % (a[0], b[0], 1e3 * (b[1] - a[1]) / (b[0] - a[0])) for a, b in zip(el[1:-1], el[2:])))
for tag in ("Loss/Total_Loss", "Loss/Loss_radiance_render", "Loss/Loss_render", "Loss/Loss_prior_albedo", "Loss/Loss_irradiance_reg"):
    print("%-28s" % tag, " ".join("%d:%.4f" % (e.step, e.value) for e in acc.Scalars(tag)))
print("checkpoints:", sorted(os.path.basename(p) for p in glob.glob(os.path.join(exp, "*.tar"))))
pngs = glob.glob(os.path.join(exp, "testset_000100", "*.png"))
print("train.py test-set export: %d PNGs, maps: %s" % (len(pngs), sorted({os.path.basename(p).rsplit("_", 1)[0] for p in pngs})))
ev = glob.glob(os.path.join(T, "logs_eval", "**", "*.png"), recursive=True)
print("test.py export: %d PNGs under logs_eval/" % len(ev))
PY
