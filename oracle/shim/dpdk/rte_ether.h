#pragma once
#include <stdint.h>
#define ETHER_ADDR_LEN 6
struct ether_addr { uint8_t addr_bytes[ETHER_ADDR_LEN]; } __attribute__((__packed__));
