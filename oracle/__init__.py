"""oracle/ — TEST INFRASTRUCTURE for the B200 dual-prompt scoring path.

A CPU restatement (`restatement.py`, fp32 PyTorch/numpy) of the reference algorithm plus the tooling
that pins it to the reference's own classes (`ref_extract.py`, `make_golden.py`) and the seeded
synthetic weights / inputs (`synth.py`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it, and only as the checker — never as the thing
measured or shipped.  The product package (`lecb200`) has no import of this directory.

Parity status: PINNED against reference-generated fixtures in tests/golden/ (see restatement.py).
"""
