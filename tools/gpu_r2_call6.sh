#!/bin/bash
# round 2, GPU call 6: grid.sync cost vs number of CTAs for the persistent kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for sz in 256 512 1024; do
  for nb in 37 74 148 296 592; do
    NSOL_LSMR_BLOCKS=$nb NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 50 --iter-max 10 --reps 3 >> gpurun_out/r2c6_admm.log 2>&1
  done
  NSOL_LSMR_PATH=1 timeout 300 python tools/time_admm.py --dim 2 --size $sz --iterations 50 --iter-max 10 --reps 3 >> gpurun_out/r2c6_admm.log 2>&1
done
for nb in 74 148 296; do
  NSOL_LSMR_BLOCKS=$nb NSOL_LSMR_PATH=4 timeout 300 python tools/time_admm.py --dim 3 --size 64 --iterations 10 --iter-max 10 --reps 3 >> gpurun_out/r2c6_admm.log 2>&1
done
sed 's/px-LSMR-it\/s.*//' gpurun_out/r2c6_admm.log
cat > /tmp/pdp.py <<'PY'
import sys, os, time, ctypes as C
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np, torch
from nsol_b200 import _lib
ctx = _lib.context()
lib = ctx.lib
for n, reg, data, alpha, iters in ((256, "TV", "L2", 0.05, 100), (1024, "HUBER", "L1", 0.6, 200)):
    img = np.random.RandomState(0).rand(n, n) * 255
    desc = _lib.PdDesc()
    desc.grid = _lib.make_grid(img.shape, None, _lib.F64, 1)
    desc.reg, desc.data, desc.alg = _lib.REG[reg], _lib.DATA[data], _lib.ALG["ALG2"]
    desc.huber_gamma, desc.L2 = 0.05, 8.0
    desc.x_scale = desc.x0_scale = desc.b_scale = float(img.max())
    al = np.array([alpha]); desc.alpha = al.ctypes.data_as(_lib.c_double_p)
    plan = C.c_void_p(); ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(plan)))
    host = np.ascontiguousarray(img.reshape(-1))
    ctx.check(lib.nsol_pd_plan_reset_host(plan, host.ctypes.data, None, None))
    for _ in range(2): ctx.check(lib.nsol_pd_plan_iterate(plan, iters, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.check(lib.nsol_pd_plan_iterate(plan, iters, None))
    e1.record(); torch.cuda.synchronize()
    print("pd persist=%s blocks=%s %dx%d: %.2f us per iteration" % (os.environ.get("NSOL_PD_PERSIST", "auto"), os.environ.get("NSOL_PD_PERSIST_BLOCKS", "auto"), n, n, e0.elapsed_time(e1) * 1e3 / (5 * iters)), flush=True)
    lib.nsol_pd_plan_destroy(plan)
PY
for nb in 16 37 74 148 296 592; do NSOL_PD_PERSIST_BLOCKS=$nb python /tmp/pdp.py >> gpurun_out/r2c6_pd.log 2>&1; done
NSOL_PD_PERSIST=2 python /tmp/pdp.py >> gpurun_out/r2c6_pd.log 2>&1
cat gpurun_out/r2c6_pd.log
