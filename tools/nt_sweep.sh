cd $GRAFT_REPO_ROOT
L=gpurun_out/s3_e2e_nt.log
: > $L
for n in 1 2 4 8; do SLS_HOST_COPY_THREADS=$n python tools/e2e_trace.py >> $L 2>&1; SLS_HOST_COPY_NO_NT=1 SLS_HOST_COPY_THREADS=$n python tools/e2e_trace.py >> $L 2>&1; done
