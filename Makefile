# Convenience targets (the driver uses __graft_entry__.py, pytest and bench.py directly).
PY ?= python

build:            ## compile libah_b200.so (nvcc, sm_100a) and the C oracle (gcc)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## oracle vs the reference's golden trajectories, C-ABI symbols, host logic, gloo world_size 2
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity of the CUDA path on a B200 (run on the GPU box)
	$(PY) -m pytest tests -q -m gpu

bench:            ## one JSON line: env-steps/s, roofline, e2e, cpu_baseline, extras
	$(PY) bench.py

golden:           ## regenerate tests/golden from the LIVE Python reference (development container only)
	$(PY) -m oracle.gen_golden

.PHONY: build test-cpu test-gpu bench golden
