for n in 1 2 4; do PROBE_UP_STREAMS=$n timeout 200 python scripts/probe_wc_upload.py 2>&1 | grep '"pinned"' | grep windows; done
