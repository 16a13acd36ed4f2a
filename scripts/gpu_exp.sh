for x in 0 1 2 3; do echo "== experiment $x"; VFI_DCN_EXPERIMENT=$x timeout 300 python scripts/dcn_debug.py 2>&1 | grep -v Warning | grep "producers\|mma \|tiles per"; done
