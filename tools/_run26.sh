set -x
SYNTH=1 timeout 240 python tools/knn_time.py 1000000 384 tc > gpurun_out/knn_time_1m_synth.log 2>&1; echo knn_exit=$?
