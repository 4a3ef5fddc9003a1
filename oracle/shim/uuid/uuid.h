/* Minimal libuuid prototypes (libuuid-dev is absent in this image); bodies in shim/uuid_shim.c.
 * Used only by the oracle/_ref build of the reference's lib/util/uuid.c. */
#pragma once
typedef unsigned char uuid_t[16];
int  uuid_parse(const char *in, uuid_t uu);
void uuid_unparse_lower(const uuid_t uu, char *out);
void uuid_generate(uuid_t out);
void uuid_copy(uuid_t dst, const uuid_t src);
int  uuid_compare(const uuid_t uu1, const uuid_t uu2);
