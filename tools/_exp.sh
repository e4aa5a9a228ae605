K="n=64 k=576 taps=9"
for d in 0 1 2 4 3 6; do EXTDM_GEMM_DBG=$d EXTDM_NO_HALO=1 python tools/gemm_experiment.py "$K" 32 2>&1 | tail -1; done
for d in 0 1 2 4 3 6; do EXTDM_GEMM_DBG=$d EXTDM_HALO_ALL=1 python tools/gemm_experiment.py "$K" 32 2>&1 | tail -1; done
EXTDM_NO_HALO=1 python tools/gemm_experiment.py "$K" 8 2>&1 | tail -1
EXTDM_HALO_ALL=1 python tools/gemm_experiment.py "$K" 8 2>&1 | tail -1
