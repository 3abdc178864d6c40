set -x
mkdir -p gpurun_out/s2
cap() { # name regex skip cmd...
  name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/s2/${name}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/s2/prof_$name -f "$@" > gpurun_out/s2/${name}_ncu.log 2>&1
  echo "$name rc=$?"
}
cap resblock5 resblock_tc_kernel 2 python scripts/run_kernel_once.py resblock5
cap coupling_f8 coupling_f8_kernel 3 python scripts/run_kernel_once.py coupling_f8
cap stencil3d stencil3d_tc_kernel 2 python scripts/run_stencil_once.py 48
ls -la gpurun_out/s2
