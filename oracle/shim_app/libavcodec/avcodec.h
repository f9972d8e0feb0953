/* TEST INFRASTRUCTURE ONLY -- the FFmpeg names App::loadAudioFile uses, as stubs that fail to open
 * anything (decoding is out of scope; inputs are synthetic).  Included inside extern "C". */
#pragma once
#include <stdint.h>
enum AVMediaType { AVMEDIA_TYPE_UNKNOWN = -1, AVMEDIA_TYPE_VIDEO, AVMEDIA_TYPE_AUDIO };
enum AVSampleFormat { AV_SAMPLE_FMT_NONE = -1, AV_SAMPLE_FMT_U8, AV_SAMPLE_FMT_S16, AV_SAMPLE_FMT_S32, AV_SAMPLE_FMT_FLT };
#define AV_CH_LAYOUT_MONO 0x4ULL
typedef struct AVCodec AVCodec;
typedef struct AVCodecParameters { enum AVMediaType codec_type; } AVCodecParameters;
typedef struct AVCodecContext {
  int codec_id, channels, sample_rate;
  uint64_t channel_layout;
  enum AVSampleFormat sample_fmt;
} AVCodecContext;
typedef struct AVPacket { int stream_index; } AVPacket;
typedef struct AVFrame { int nb_samples; uint8_t *data[8]; } AVFrame;
static inline const AVCodec *avcodec_find_decoder(int id) { (void)id; return 0; }
static inline int avcodec_open2(AVCodecContext *c, const AVCodec *d, void *o) { (void)c; (void)d; (void)o; return -1; }
static inline int avcodec_close(AVCodecContext *c) { (void)c; return 0; }
static inline int avcodec_decode_audio4(AVCodecContext *c, AVFrame *f, int *got, const AVPacket *p) { (void)c; (void)f; (void)p; *got = 0; return -1; }
static inline AVPacket *av_packet_alloc(void) { return 0; }
static inline void av_packet_unref(AVPacket *p) { (void)p; }
static inline AVFrame *av_frame_alloc(void) { return 0; }
static inline void av_frame_free(AVFrame **f) { (void)f; }
static inline int av_samples_alloc(uint8_t **b, int *l, int ch, int n, enum AVSampleFormat f, int a) { (void)b; (void)l; (void)ch; (void)n; (void)f; (void)a; return -1; }
