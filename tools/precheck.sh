#!/usr/bin/env bash
# CPU-side gate to run before every gpurun call: syntax, imports, ABI symbols.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import ast, glob, sys
sys.path.insert(0, '.')
for f in glob.glob('qpgesture_b200/*.py') + glob.glob('tests/*.py') + glob.glob('tools/*.py') + ['bench.py', '__graft_entry__.py'] + glob.glob('oracle/*.py'):
    ast.parse(open(f).read())
import qpgesture_b200.matchdb, qpgesture_b200.GestureKNN, qpgesture_b200.vqvae, qpgesture_b200.VisualizeCodebook, qpgesture_b200.sharding, qpgesture_b200.PAE, qpgesture_b200.process_bvh
from qpgesture_b200 import _lib
_lib.load()
print("precheck ok")
PY
