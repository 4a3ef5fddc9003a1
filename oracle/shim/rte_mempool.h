/* empty: DPDK header stand-in for the oracle/_ref build */
