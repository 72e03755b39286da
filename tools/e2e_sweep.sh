# run on the GPU box: the host drop-in call (tools/e2e_trace.py) under different chunkings and rowval sources; log to gpurun_out/
cd $GRAFT_REPO_ROOT
L=gpurun_out/${E2E_LOG:-e2e_sweep.log}
: > $L
SLS_PIPE_TRACE=1 python tools/e2e_trace.py >> $L 2>&1
SLS_PIPE_ROWS_FROM_DEVICE=1 SLS_PIPE_TRACE=1 python tools/e2e_trace.py >> $L 2>&1
for n in 1 2 4 8 16; do SLS_HOST_COPY_THREADS=$n python tools/e2e_trace.py >> $L 2>&1; done
for k in 2 3 5 6 8; do SLS_PIPE_CHUNKS=$k python tools/e2e_trace.py >> $L 2>&1; done
for f in "2,3,3,2" "1,2,3,4" "1,1.6,2.5,4,3"; do SLS_PIPE_FRACS=$f python tools/e2e_trace.py >> $L 2>&1; done
python tools/e2e_trace.py --values-only >> $L 2>&1
