for p in 1 2 3; do
  echo "=== path $p"
  OCRPP_DEBUG_SYNC=1 ONLY_PATH=$p python tests/dev_db_determinism.py 40 2>&1 | grep -v "^frame\|^<omit\|^Search\|^CUDA kernel\|^For debug\|^Compile with" | tail -8
done
