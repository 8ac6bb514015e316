cd "$(dirname "$0")"
run() { echo "--- $*"; env "$@" ./t3_timeline $E $N | sed -n '1,4p'; }
E=4096; N=4
for LR in 1 2 4; do run QRL_EM_T3_LR=$LR QRL_EM_T3_C=24 QRL_EM_T3_D=4; done
run QRL_EM_T3_LR=2 QRL_EM_T3_C=12 QRL_EM_T3_D=4
run QRL_EM_T3_LR=2 QRL_EM_T3_C=36 QRL_EM_T3_D=3
run QRL_EM_T3_LR=2 QRL_EM_T3_SPLIT=1 QRL_EM_T3_C=24 QRL_EM_T3_D=4
E=8192; N=8
for LR in 4 8; do run QRL_EM_T3_LR=$LR QRL_EM_T3_C=8 QRL_EM_T3_D=4; done
