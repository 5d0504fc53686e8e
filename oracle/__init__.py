"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's (YoungSeng/QPGesture @ 7dd5daa) inference
hot path, used as the checker in tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py.  Nothing under
qpgesture_b200/ imports this package: the product path is CUDA-only and
raises if the native library is missing.

Pinning status (see DESIGN.md "Oracle"): the reference ships no tests, golden
vectors or fixtures, so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, imported in place from /root/reference by oracle/ref_harness.py in
the build container (it cannot travel to the GPU box); the generated vectors
are committed under tests/golden/ together with tests/golden/make_golden.py.
"""
