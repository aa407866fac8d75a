#!/bin/bash
# compute-sanitizer over the kernels with cross-CTA / cross-GPU protocols, on small shapes (ONE tool per gpurun call):
#   bash tools/gpu_sanitizer.sh memcheck|racecheck|synccheck
# - in-kernel z-slab halo exchange (link flags, peer stores; single-GPU emulation of 2 / 3 slabs)
# - cooperative LSMR solves (lsmr_coop, lsmr_coopv) and the persistent primal-dual kernel (grid.sync)
# - fused 2-D / 3-D LSMR kernels (shared-memory rings, cp.async staging) and the TMA bulk primal-dual kernel
TOOL=${1:-memcheck}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SEL='zslab_link_emulation_matches_unsharded or zslab_link_timeout or lsmr_multi_kernel_and_cooperative_paths or lsmr_fused_2d_kernels_match_generic_kernels or lsmr_fused_3d_kernels_match_generic_kernels or primal_dual_tiling_and_variant_independent or primal_dual_ragged_shapes_vs_oracle or admm_with_b_reg or nan_and_inf'
# the plain run must pass first
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "$SEL" 2>&1 | tail -3 > gpurun_out/san_${TOOL}_plain.log
cat gpurun_out/san_${TOOL}_plain.log
grep -q " passed" gpurun_out/san_${TOOL}_plain.log && ! grep -q "failed" gpurun_out/san_${TOOL}_plain.log || exit 1
timeout 2400 compute-sanitizer --tool $TOOL --error-exitcode 77 --print-limit 20 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "$SEL" > gpurun_out/san_${TOOL}.log 2>&1
echo "sanitizer exit: $?" >> gpurun_out/san_${TOOL}.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|sanitizer exit" gpurun_out/san_${TOOL}.log | tail -8
