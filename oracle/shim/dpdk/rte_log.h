#pragma once
#include <stdio.h>
#define RTE_LOGTYPE_USER1 24
#define RTE_LOG_ERR 4U
#define RTE_LOG_WARNING 5U
#define RTE_LOG_NOTICE 6U
#define RTE_LOG_INFO 7U
#define RTE_LOG_DEBUG 8U
int rte_log(uint32_t level, uint32_t logtype, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
#define RTE_LOG(l, t, ...) rte_log(RTE_LOG_##l, RTE_LOGTYPE_##t, #t ": " __VA_ARGS__)
