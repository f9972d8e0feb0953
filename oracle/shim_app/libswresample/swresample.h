/* TEST INFRASTRUCTURE ONLY -- see libavcodec/avcodec.h. */
#pragma once
#include <stdint.h>
struct SwrContext;
static inline struct SwrContext *swr_alloc(void) { return 0; }
static inline int swr_init(struct SwrContext *s) { (void)s; return -1; }
static inline int swr_is_initialized(struct SwrContext *s) { (void)s; return 0; }
static inline int swr_convert(struct SwrContext *s, uint8_t **o, int oc, const uint8_t **i, int ic) { (void)s; (void)o; (void)oc; (void)i; (void)ic; return 0; }
static inline void swr_free(struct SwrContext **s) { (void)s; }
