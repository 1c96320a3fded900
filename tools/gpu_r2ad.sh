#!/bin/bash
# Final validation of the round: full GPU suite, smoke, the driver's bench line and the reference arm on one GPU.
O=gpurun_out
TAG=${1:-r02d}
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log )
tail -3 $O/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
S=$(date +%s); timeout 900 python bench.py > $O/b_default_$TAG.json 2> $O/b_default_$TAG.err; echo "bench default exit $? wall $(( $(date +%s) - S )) s"
S=$(date +%s); timeout 900 python bench.py --impl reference > $O/b_ref_$TAG.json 2> $O/b_ref_$TAG.err; echo "bench ref exit $? wall $(( $(date +%s) - S )) s"
python tools/summ.py default_$TAG 2>&1 | tee $O/summ_$TAG.txt
