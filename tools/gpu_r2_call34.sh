#!/bin/bash
# round 2, GPU call 34: ncu --set full of the link=1 iteration kernel on a 64-plane slab (single-GPU emulation of two linked slabs) and
# of the link=0 kernel on the same slab
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/link_emulation_profile.py > gpurun_out/r2c34_plain.log 2>&1 || { tail -5 gpurun_out/r2c34_plain.log; exit 1; }
timeout 300 python tools/link_emulation_profile.py --unlinked >> gpurun_out/r2c34_plain.log 2>&1
cat gpurun_out/r2c34_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pd_iter_bulk -s 8 -c 2 -o gpurun_out/r2c34_link1 -f python tools/link_emulation_profile.py > gpurun_out/r2c34_ncu1.log 2>&1; echo "link=1 capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pd_iter_bulk -s 8 -c 2 -o gpurun_out/r2c34_link0 -f python tools/link_emulation_profile.py --unlinked > gpurun_out/r2c34_ncu0.log 2>&1; echo "link=0 capture rc=$?"
ls -la gpurun_out/r2c34*.ncu-rep
