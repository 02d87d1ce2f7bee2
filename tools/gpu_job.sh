#!/bin/bash
# Round-end check in one gpurun call: the GPU parity suite, the smoke entry and the bench line of both arms.
# (During development this file was rewritten per experiment; the durable tools are tools/run_op.py, tools/e2e_sweep.py and bench.py.)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-200
python bench.py --steps 50 --warmup 5 --no-ops --no-cpu 2>gpurun_out/b.err | cut -c1-400
