#!/bin/bash
# ncu --set full captures of the four fast strip kernels and the two deferred-lines kernels at 2048^2 (one launch each, taken
# after a complete update so that the state is realistic); raw pages as csv + the reports (source page) under gpurun_out/
#   scripts/gpu_session_ncu.sh tag        -> gpurun_out/<tag>_{strip,lines}_{mevp,bbm}_{rect,para}.{csv,ncu-rep}
cd "$(dirname "$0")/.."
tag=${1:-ncu}
mkdir -p gpurun_out
cap() { # name kernel-regex env...
  local name=$1 kre=$2; shift 2
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s 30 -c 1 -f -o gpurun_out/${tag}_$name \
    python scripts/quickbench.py > gpurun_out/${tag}_$name.log 2>&1
  ncu -i gpurun_out/${tag}_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_$name.csv 2>/dev/null
  python scripts/ncu_csv_summary.py gpurun_out/${tag}_$name.csv gpurun_out/${tag}_$name.txt | grep -E "^## |gpu__time_duration|dram__bytes|registers_per_thread" 
}
cap strip_mevp_rect subcycle_strip QB_RHEO=mevp
cap strip_bbm_rect  subcycle_strip QB_RHEO=bbm
cap strip_mevp_para subcycle_strip QB_RHEO=mevp QB_DISTORT=1
cap strip_bbm_para  subcycle_strip QB_RHEO=bbm QB_DISTORT=1
cap lines_mevp_rect subcycle_lines QB_RHEO=mevp
cap lines_bbm_rect  subcycle_lines QB_RHEO=bbm
du -sh gpurun_out
