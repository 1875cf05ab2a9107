for cw in "1024 192" "1024 512" "1024 1024" "2048 512" "2048 1024" "512 512"; do set -- $cw
echo "chunk $1 warm $2"; CHUNK=$1 WARM=$2 python scripts/diag_cfg4.py 5 2>&1 | grep -E "^[0-9] estep|estep_ms|chunk runs" | tail -9
done
