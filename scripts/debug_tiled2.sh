for m in 1 2 0; do echo "== MCS_DEBUG_TILED=$m"; MCS_DEBUG_TILED=$m CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/debug_tiled.py 2>&1 | tail -4; done
