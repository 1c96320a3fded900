#!/bin/bash
# Round-2 call G: full GPU suite, smoke, default bench (wall-clocked), reference arm.
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_r2g.log 2>&1; echo "pytest exit $?" >> $O/pytest_r2g.log )
tail -3 $O/pytest_r2g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2g.log 2>&1; tail -2 $O/smoke_r2g.log
S=$(date +%s); timeout 1200 python bench.py > $O/b_default_r2g.json 2> $O/b_default_r2g.err; echo "bench default exit $? wall $(( $(date +%s) - S )) s"
S=$(date +%s); timeout 900 python bench.py --impl reference > $O/b_ref_r2g.json 2> $O/b_ref_r2g.err; echo "bench ref exit $? wall $(( $(date +%s) - S )) s"
tail -c 3000 $O/b_default_r2g.json
