/* Stand-in for DPDK's rte_config.h: the three macros lib/vhost needs (oracle/_ref build only). */
#pragma once
#define RTE_CACHE_LINE_SIZE 64
#define __rte_cache_aligned __attribute__((aligned(RTE_CACHE_LINE_SIZE)))
#define __rte_unused __attribute__((unused))
#ifndef likely
#define likely(x)   __builtin_expect(!!(x), 1)
#endif
#ifndef unlikely
#define unlikely(x) __builtin_expect(!!(x), 0)
#endif
