# convenience targets; the driver calls __graft_entry__.build() / pytest / bench.py directly
PY ?= python

build:            ## liboimgpu.so (nvcc, sm_100a), oim-gpu-vhost, the CPU checkers
	$(PY) __graft_entry__.py

test: build       ## everything that runs without a GPU
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity tests proper (B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) __graft_entry__.py smoke

bench: build
	$(PY) bench.py --steps 10 --warmup 3

campaign: build   ## restatement vs reference and wire protocols on many more seeds (CPU)
	$(PY) tests/campaign_oracle.py
	$(PY) tests/campaign_transport.py

.PHONY: build test test-gpu smoke bench campaign
