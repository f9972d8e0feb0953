/* TEST INFRASTRUCTURE ONLY -- see libavcodec/avcodec.h. */
#pragma once
#include <libavcodec/avcodec.h>
typedef struct AVStream { AVCodecParameters *codecpar; AVCodecContext *codec; } AVStream;
typedef struct AVFormatContext { unsigned nb_streams; AVStream **streams; } AVFormatContext;
static inline AVFormatContext *avformat_alloc_context(void) { return 0; }
static inline int avformat_open_input(AVFormatContext **c, const char *u, void *f, void *o) { (void)c; (void)u; (void)f; (void)o; return -1; }
static inline int avformat_find_stream_info(AVFormatContext *c, void *o) { (void)c; (void)o; return -1; }
static inline void avformat_free_context(AVFormatContext *c) { (void)c; }
static inline int av_read_frame(AVFormatContext *c, AVPacket *p) { (void)c; (void)p; return -1; }
