# run on the GPU box: the host drop-in call under different chunkings (tools/e2e_trace.py), traces to gpurun_out/
cd $GRAFT_REPO_ROOT
L=gpurun_out/${E2E_LOG:-e2e_sweep.log}
: > $L
SLS_PIPE_TRACE=1 python tools/e2e_trace.py >> $L 2>&1
SLS_PIPE_NZ_WHOLE=1 SLS_PIPE_TRACE=1 python tools/e2e_trace.py >> $L 2>&1
for f in "2,3,3,2" "1.5,3,3.5,2" "1,2,3,3,1" "1,3,4,2" "1,2,3,4"; do SLS_PIPE_TRACE=1 SLS_PIPE_FRACS=$f python tools/e2e_trace.py >> $L 2>&1; done
for k in 3 5; do SLS_PIPE_CHUNKS=$k python tools/e2e_trace.py >> $L 2>&1; done
python tools/e2e_trace.py --values-only >> $L 2>&1
