/*
 * ref_nut.c — TEST INFRASTRUCTURE ONLY.  FATE's filter-pixfmts-* goldens are md5 sums of a NUT stream holding one
 * rawvideo frame (tests/fate-run.sh:620-659: `ffmpeg ... -vf "scale,format=FMT,FILTER" -vcodec rawvideo -pix_fmt FMT
 * -frames:v 1 -f nut md5:` with -flags +bitexact -fflags +bitexact).  This file drives the UNMODIFIED NUT muxer of the
 * reference (libavformat/nutenc.c, compiled where it lies by oracle/ref/Makefile) the way the ffmpeg tool does for that
 * command, so that a frame produced by the oracle or by the CUDA path can be checked against the md5 sums committed in
 * the reference tree (tests/ref/fate/filter-pixfmts-null, -copy, -scale).
 *
 * What the tool sets on the output stream (fftools/ffmpeg_mux_init.c, libavcodec/rawenc.c):
 *   codec rawvideo, codec_tag = avcodec_pix_fmt_to_codec_tag(pix_fmt) (rawenc.c raw_encode_init), time base 1/25 (the image2
 *   demuxer's default frame rate), avg_frame_rate 25/1, stream metadata encoder = "Lavc rawvideo" (set_encoder_id() with
 *   -flags +bitexact, ffmpeg_mux_init.c:1113-1133), no global metadata with -fflags +bitexact.
 *
 * The symbols at the bottom are the parts of libavcodec / libavformat the muxing path links against but never reaches
 * for one rawvideo stream (bitstream filters, protocol layer, format registry, decoder lookup).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libavformat/avformat.h"
#include "libavformat/mux.h"
#include "libavformat/avio_internal.h"
#include "libavutil/md5.h"
#include "libavutil/opt.h"
#include "libavutil/mem.h"
#include "libavutil/dict.h"
#include "libavutil/log.h"
#include "libavutil/pixdesc.h"
#include "libavutil/mathematics.h"
#include "libavcodec/packet.h"
#include "libavcodec/codec_par.h"
#include "libavcodec/raw.h"
#include "libavcodec/avcodec.h"

#define API __attribute__((visibility("default")))

extern const FFOutputFormat ff_nut_muxer;

/* frames: nframes rawvideo packets of `size` bytes back to back (each: planes back to back, line size = width * bytes
 * per pixel), pts 0,1,2...; returns the NUT stream size.  nframes = 1 is the filter-pixfmts-* command, nframes = 5 the
 * filter-pixdesc-* one (tests/fate-run.sh:599-618, `-frames:v 5`). */
API int ffref_nut_md5_frames(const uint8_t *frames, int size, int nframes, int w, int h, int pix_fmt, uint8_t md5[16])
{
    AVFormatContext *s = avformat_alloc_context();
    AVPacket *pkt = NULL;
    uint8_t *buf = NULL;
    int ret = -1, n;
    if (!s) return -1;
    s->oformat = &ff_nut_muxer.p;
    if (ff_nut_muxer.priv_data_size) {
        s->priv_data = av_mallocz(ff_nut_muxer.priv_data_size);
        if (!s->priv_data) goto end;
        if (ff_nut_muxer.p.priv_class) {
            *(const AVClass **)s->priv_data = ff_nut_muxer.p.priv_class;
            av_opt_set_defaults(s->priv_data);
        }
    }
    s->flags |= AVFMT_FLAG_BITEXACT;
    AVStream *st = avformat_new_stream(s, NULL);
    if (!st) goto end;
    st->codecpar->codec_type = AVMEDIA_TYPE_VIDEO;
    st->codecpar->codec_id   = AV_CODEC_ID_RAWVIDEO;
    st->codecpar->format     = pix_fmt;
    st->codecpar->width      = w;
    st->codecpar->height     = h;
    st->codecpar->codec_tag  = avcodec_pix_fmt_to_codec_tag(pix_fmt);
    st->codecpar->bits_per_coded_sample = av_get_bits_per_pixel(av_pix_fmt_desc_get(pix_fmt));
    st->time_base      = (AVRational){ 1, 25 };
    st->avg_frame_rate = (AVRational){ 25, 1 };
    av_dict_set(&st->metadata, "encoder", "Lavc rawvideo", 0);
    if (avio_open_dyn_buf(&s->pb) < 0) goto end;
    if ((ret = avformat_write_header(s, NULL)) < 0) goto end;
    pkt = av_packet_alloc();
    if (!pkt) { ret = -1; goto end; }
    for (int i = 0; i < nframes; i++) {
        if (av_new_packet(pkt, size) < 0) { ret = -1; goto end; }
        memcpy(pkt->data, frames + (size_t)i * size, size);
        /* the muxer picked its own stream time base in write_header (nutenc.c:753-756); the tool rescales the
         * encoder's 1/25 timestamps to it (fftools/ffmpeg_mux.c) */
        pkt->pts = pkt->dts = av_rescale_q(i, (AVRational){ 1, 25 }, st->time_base);
        pkt->duration = av_rescale_q(1, (AVRational){ 1, 25 }, st->time_base);
        pkt->flags |= AV_PKT_FLAG_KEY;
        pkt->stream_index = 0;
        if ((ret = av_write_frame(s, pkt)) < 0) goto end;
        av_packet_unref(pkt);
    }
    av_write_trailer(s);
    n = avio_close_dyn_buf(s->pb, &buf);
    s->pb = NULL;
    av_md5_sum(md5, buf, n);
    ret = n;
end:
    av_free(buf);
    av_packet_free(&pkt);
    if (s && s->pb) { uint8_t *b2 = NULL; avio_close_dyn_buf(s->pb, &b2); av_free(b2); s->pb = NULL; }
    avformat_free_context(s);
    return ret;
}

API int ffref_nut_md5(const uint8_t *frame, int size, int w, int h, int pix_fmt, uint8_t md5[16])
{
    return ffref_nut_md5_frames(frame, size, 1, w, h, pix_fmt, md5);
}

/* ------------------------------------------------------------------ never reached for one rawvideo stream */
struct AVBSFContext; struct AVBitStreamFilter; struct AVCodecParserContext;
void av_bsf_free(struct AVBSFContext **c) { if (c) *c = NULL; }
int av_bsf_init(struct AVBSFContext *c) { return -1; }
int av_bsf_send_packet(struct AVBSFContext *c, AVPacket *p) { return -1; }
int av_bsf_receive_packet(struct AVBSFContext *c, AVPacket *p) { return -1; }
const struct AVBitStreamFilter *av_bsf_get_by_name(const char *n) { return NULL; }
int av_bsf_alloc(const struct AVBitStreamFilter *f, struct AVBSFContext **c) { return -1; }
const AVOutputFormat *av_guess_format(const char *a, const char *b, const char *c) { return NULL; }
int av_get_audio_frame_duration2(AVCodecParameters *par, int frame_bytes) { return 0; }
const char *avcodec_get_name(enum AVCodecID id) { return "unknown"; }
void avcodec_free_context(AVCodecContext **c) { if (c) { av_free(*c); *c = NULL; } }
const AVCodec *avcodec_find_decoder(enum AVCodecID id) { return NULL; }
void av_parser_close(struct AVCodecParserContext *s) { }
int av_get_bits_per_sample(enum AVCodecID codec_id) { return 0; }
const AVClass ff_avio_class = { .class_name = "AVIOContext", .item_name = av_default_item_name, .version = LIBAVUTIL_VERSION_INT };
int ffio_open_whitelist2(AVIOContext **s, const char *url, int flags, const AVIOInterruptCB *int_cb, AVDictionary **options,
                         const char *whitelist, const char *blacklist, AVFormatContext *avfc) { return -1; }
const char *avio_find_protocol_name(const char *url) { return NULL; }
int avio_close(AVIOContext *s) { return 0; }
const AVOutputFormat *av_muxer_iterate(void **opaque) { return NULL; }
const AVInputFormat *av_demuxer_iterate(void **opaque) { return NULL; }
/* seek.c (the stream index the muxer keeps for syncpoint back pointers) reaches these only when demuxing; a muxed stream
 * has pts_wrap_behavior = AV_PTS_WRAP_IGNORE, for which libavformat/demux.c:52-54 returns the timestamp unchanged */
int64_t ff_wrap_timestamp(const AVStream *st, int64_t timestamp) { return timestamp; }
int av_read_frame(AVFormatContext *s, AVPacket *pkt) { return -1; }
int avformat_queue_attached_pictures(AVFormatContext *s) { return -1; }
/* avformat_new_stream() wants an internal codec context; the NUT muxing path never looks inside it */
AVCodecContext *avcodec_alloc_context3(const AVCodec *c) { return av_mallocz(16384); }
void ff_parse_specific_params(AVStream *st, int *au_rate, int *au_ssize, int *au_scale) { *au_rate = 1; *au_ssize = 1; *au_scale = 1; }
