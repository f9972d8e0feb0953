/* TEST INFRASTRUCTURE ONLY -- see libavcodec/avcodec.h. */
#pragma once
#include <libavcodec/avcodec.h>
static inline int av_opt_set_int(void *o, const char *n, int64_t v, int f) { (void)o; (void)n; (void)v; (void)f; return 0; }
static inline int av_opt_set_sample_fmt(void *o, const char *n, enum AVSampleFormat v, int f) { (void)o; (void)n; (void)v; (void)f; return 0; }
