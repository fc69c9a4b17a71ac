# Runs the shipped example visit (examples/hd209458b_like, 121 exposures) three times on the GPU:
# cold, steady state without FITS output, steady state with FITS output.  usage: bash tools/run_example_visit.sh
set -e
export WAYNE_CALB_DIR=/tmp/calb_example
python -c "from wayne_b200 import calibration; calibration.write_synthetic_calibration('/tmp/calb_example')" >/dev/null
cp -r examples/hd209458b_like /tmp/visit_ex
python - <<'P'
import time, yaml, os, sys
sys.path.insert(0, os.getcwd())
from wayne_b200 import run_visit
import torch
p='/tmp/visit_ex/params.yml'
cfg=yaml.safe_load(open(p))
# warm-up: a first visit (library load, contexts, pinned pools)
t0=time.perf_counter(); out=run_visit.run(['-p', p]); torch.cuda.synchronize(); t1=time.perf_counter()
print('first visit (cold):', round(t1-t0,2),'s', len(out),'exposures')
obs=run_visit.build_observation(cfg, '/tmp/visit_ex')
t0=time.perf_counter(); r=obs.run_observation(write_fits=False); torch.cuda.synchronize(); t1=time.perf_counter()
print('visit, no FITS:', round(t1-t0,3),'s', len(r))
obs=run_visit.build_observation(cfg, '/tmp/visit_ex')
t0=time.perf_counter(); r=obs.run_observation(write_fits=True); torch.cuda.synchronize(); t1=time.perf_counter()
print('visit, with FITS:', round(t1-t0,3),'s', len(r))
P
