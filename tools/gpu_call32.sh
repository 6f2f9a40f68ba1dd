timeout 120 ./tools/tma_prior_bench 384 16384; echo rc=$?
timeout 120 ./tools/tma_prior_bench 64 16384; echo rc=$?
