set -x
mkdir -p gpurun_out; rm -f gpurun_out/mp_ab.txt
timeout 600 python tools/mp_ab.py QAGNN_MP_SPLIT=0 QAGNN_MP_SPLIT=1 QAGNN_MP_SPLIT=1,QAGNN_MP_WARPS2=28 QAGNN_MP_SPLIT=1,QAGNN_MP_WARPS2=31 QAGNN_MP_SPLIT=1,QAGNN_MP_WARPS=31,QAGNN_MP_WARPS2=31 QAGNN_MP_SPLIT=1,QAGNN_MP_WARPS=28,QAGNN_MP_WARPS2=31 > gpurun_out/r2v_ab.log 2>&1
grep -E "^cfg2 |^cfg2-loader" gpurun_out/mp_ab.txt
