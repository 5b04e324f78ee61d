# K1 probe across claim-batch sizes and profiling modes (0 product, 1 stream only, 2 no atomics, 3 payload touch)
for b in ${BATCHES:-2 4 8}; do for m in ${MODES:-0 1 3}; do echo "batch $b mode $m"; SILO_K1_BATCH=$b SILO_K1_STREAM_ONLY=$m python profiles/k1_probe.py 2>/dev/null | head -${LINES_PER_RUN:-2}; done; done
