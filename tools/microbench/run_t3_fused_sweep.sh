# Sweep of the fused (generator interleaved with the piece pass) worker loop of em_tile3_kernel against the unfused one:
# chunk length C x ring depth D x pairs per trip, pure device time of 200 back-to-back launches.
cd "$(dirname "$0")"
run() { bin=$1; shift; printf "%-14s %-40s " "$bin" "$*"; env "$@" ./$bin $E $N | head -1; }
for cfg in "4096 4" "8192 8" "8192 4" "65536 8"; do
  set -- $cfg; E=$1; N=$2
  echo "=== E=$E n=$N"
  run t3b_unfused QRL_EM_T3_C=36 QRL_EM_T3_D=3
  for bin in t3b_fused_p2 t3b_fused_p1 t3b_fused_p3; do
    for knobs in "QRL_EM_T3_C=36 QRL_EM_T3_D=3" "QRL_EM_T3_C=24 QRL_EM_T3_D=4" "QRL_EM_T3_C=18 QRL_EM_T3_D=6" "QRL_EM_T3_C=18 QRL_EM_T3_D=4" "QRL_EM_T3_C=12 QRL_EM_T3_D=8" "QRL_EM_T3_C=12 QRL_EM_T3_D=4" "QRL_EM_T3_C=6 QRL_EM_T3_D=8"; do
      run $bin $knobs
    done
  done
done
echo "=== timeline (instrumented), E=4096 n=4"
E=4096; N=4
for knobs in "QRL_EM_T3_C=36 QRL_EM_T3_D=3" "QRL_EM_T3_C=18 QRL_EM_T3_D=6" "QRL_EM_T3_C=12 QRL_EM_T3_D=8"; do
  echo "--- $knobs"; env T3_TRACE=1 $knobs ./t3_timeline $E $N
done
