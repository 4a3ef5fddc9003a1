/* test-build stand-in for DPDK's generated rte_config.h, force-included as S/lib/vhost/rte_vhost/Makefile
 * does; DPDK's own headers pull these in transitively */
#pragma once
#include <limits.h>
#include <inttypes.h>
#include <linux/limits.h>
#include <errno.h>
#include "rte_common.h"
#define RTE_MAX_ETHPORTS 32
