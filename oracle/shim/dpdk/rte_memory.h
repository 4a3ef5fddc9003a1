#pragma once
#include <stdint.h>
typedef uint64_t phys_addr_t;
phys_addr_t rte_mem_virt2phy(const void *virt);
