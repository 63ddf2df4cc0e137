#!/bin/bash
# session r4a: ranged refit walk — the refit GPU tests and the timing of one moved entity (ranged vs whole-tree walk)
OUT=gpurun_out/r4a; mkdir -p $OUT
(time timeout 300 python -m pytest tests -m gpu -x -q -k "refit") > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
timeout 200 python tools/refit_time.py > $OUT/refit_time.json 2> $OUT/refit_time.err
tail -5 $OUT/pytest.log; cat $OUT/refit_time.json; tail -3 $OUT/refit_time.err
