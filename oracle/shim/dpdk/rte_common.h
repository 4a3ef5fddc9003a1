/* stand-ins for the few DPDK helpers S/lib/vhost/rte_vhost uses; written against their documented meaning */
#pragma once
#include <stdint.h>
#include <stddef.h>
#define RTE_MIN(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a < _b ? _a : _b; })
#define RTE_MAX(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a > _b ? _a : _b; })
#define RTE_ALIGN_FLOOR(v, a) ((__typeof__(v))((v) & ~((__typeof__(v))((a) - 1))))
#define RTE_ALIGN_CEIL(v, a) RTE_ALIGN_FLOOR(((v) + ((__typeof__(v))(a) - 1)), a)
#define RTE_SET_USED(x) (void)(x)
#define RTE_PTR_ADD(p, x) ((void *)((uintptr_t)(p) + (x)))
#define RTE_CACHE_LINE_SIZE 64
#define __rte_cache_aligned __attribute__((aligned(RTE_CACHE_LINE_SIZE)))
#define __rte_unused __attribute__((unused))
#define __rte_experimental
#define likely(x) __builtin_expect(!!(x), 1)
#define unlikely(x) __builtin_expect(!!(x), 0)
#define rte_smp_wmb() __asm__ volatile("" ::: "memory")
#define rte_smp_rmb() __asm__ volatile("" ::: "memory")
#define rte_smp_mb() __sync_synchronize()
typedef struct { volatile int16_t cnt; } rte_atomic16_t;
static inline void rte_atomic16_set(rte_atomic16_t *v, int16_t n) { v->cnt = n; }
