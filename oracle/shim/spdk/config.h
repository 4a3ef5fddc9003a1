/* Stand-in for the header SPDK's ./configure generates (all CONFIG_* left undefined:
 * no DPDK, no ASAN, no debug). Only used when building oracle/_ref from /root/reference. */
#pragma once
