# the GPU suite under non-default global options: no programmatic dependent launch, PDL inside operators only, the
# multi-kernel plan build, no TMA staging, one-CTA FPS / EMD
mkdir -p gpurun_out
for o in "pdl=0" "pdl=1" "plan=1" "stage=1" "fps=1,emd_cluster=0,chamfer_once=0" "grouping_bwd=1,three_nn_bwd=1,gather_bwd=1"; do
  echo "== $o"; PCFM_B200_OPTIONS="$o" timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
done | tee gpurun_out/option_matrix.log
