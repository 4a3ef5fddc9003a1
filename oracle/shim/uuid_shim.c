/* 5 libuuid functions, enough for SPDK's uuid wrapper (test infrastructure only). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "uuid/uuid.h"

int uuid_parse(const char *in, uuid_t uu)
{
	int i, n = 0;
	if (strlen(in) != 36) return -1;
	for (i = 0; i < 36; i++) {
		if (i == 8 || i == 13 || i == 18 || i == 23) {
			if (in[i] != '-') return -1;
			continue;
		}
		if (!isxdigit((unsigned char)in[i])) return -1;
		unsigned v = isdigit((unsigned char)in[i]) ? in[i] - '0' : (tolower(in[i]) - 'a' + 10);
		if (n & 1) uu[n >> 1] |= v; else uu[n >> 1] = v << 4;
		n++;
	}
	return 0;
}
void uuid_unparse_lower(const uuid_t uu, char *out)
{
	sprintf(out, "%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x",
		uu[0], uu[1], uu[2], uu[3], uu[4], uu[5], uu[6], uu[7],
		uu[8], uu[9], uu[10], uu[11], uu[12], uu[13], uu[14], uu[15]);
}
void uuid_generate(uuid_t out)
{
	FILE *f = fopen("/dev/urandom", "rb");
	if (!f || fread(out, 1, 16, f) != 16) { for (int i = 0; i < 16; i++) out[i] = rand(); }
	if (f) fclose(f);
	out[6] = (out[6] & 0x0f) | 0x40;
	out[8] = (out[8] & 0x3f) | 0x80;
}
void uuid_copy(uuid_t dst, const uuid_t src) { memcpy(dst, src, 16); }
int  uuid_compare(const uuid_t a, const uuid_t b) { return memcmp(a, b, 16); }
