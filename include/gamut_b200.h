/*
 * gamut_b200.h -- C ABI of the Blackwell-native decode/convert engine behind the Gamut API.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and replaces one seam of the
 * reference (AuburnSounds/gamut @ 9697a371) that a betterC D shim can bind with
 * `extern(C) nothrow @nogc` (see INTEGRATION.md and d/gamut_b200.d). Reference citations are
 * `file:line` under the reference's source/gamut/.
 *
 * Conventions
 *  - "host" entry points take host pointers, return malloc()'d host pixels (free with free() /
 *    gb200_free) exactly like the reference codecs (plugins/png.d:108, image.d:27-30), and are
 *    synchronous and thread-safe.
 *  - "_device" / "_batch" entry points work on device-resident buffers on a caller-provided CUDA
 *    stream (`void* stream` is a cudaStream_t; NULL = legacy default stream) and are asynchronous.
 *  - Failure is reported like the reference codecs do: NULL / 0, with a static thread-local
 *    message retrievable by gb200_last_error() (the analogue of Image.errorMessage, image.d:390).
 *  - There is NO CPU fallback: without an sm_100 device every entry point fails loudly.
 */
#ifndef GAMUT_B200_H
#define GAMUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PixelType -- integer values are ABI and equal gamut.types.PixelType (types.d:32-59). */
enum gb200_pixel_type {
    GB200_unknown = -1,
    GB200_l8 = 0, GB200_l16 = 1, GB200_lf32 = 2,
    GB200_la8 = 3, GB200_la16 = 4, GB200_laf32 = 5,
    GB200_lap8 = 6, GB200_lap16 = 7, GB200_lapf32 = 8,
    GB200_rgb8 = 9, GB200_rgb16 = 10, GB200_rgbf32 = 11,
    GB200_rgba8 = 12, GB200_rgba16 = 13, GB200_rgbaf32 = 14,
    GB200_rgbap8 = 15, GB200_rgbap16 = 16, GB200_rgbapf32 = 17
};

/* ImageFormat (types.d:14-28), only the four formats on the hot path. */
enum gb200_image_format { GB200_FORMAT_JPEG = 0, GB200_FORMAT_PNG = 1, GB200_FORMAT_QOI = 2, GB200_FORMAT_QOIX = 3 };

/* ---- library ---- */
int         gb200_init(void);               /* 1 if an sm_100 device is usable */
const char* gb200_version(void);
const char* gb200_last_error(void);         /* thread-local, static storage, "" if none */
long long   gb200_launch_count(void);       /* kernels launched by this library so far */
int         gb200_sm_count(void);
void*       gb200_device_alloc(size_t bytes);  /* cached cudaMalloc */
void        gb200_device_free(void* p);
void        gb200_device_trim(void);
void*       gb200_host_alloc(size_t bytes);    /* pinned host memory for the host entry points */
void        gb200_host_free(void* p);
void        gb200_free(void* p);               /* free() for pixels returned by host decoders */

/* ---- PixelType converters: source/gamut/scanline.d ---- */
int gb200_pixel_type_size(int type);                       /* pixelTypeSize, types.d:62 */
int gb200_scanlines_inter_type(int srcType, int dstType);  /* scanlinesInterType, scanline.d:25 */

/* scanlinesConvert (scanline.d:70-121) and, for srcType == dstType, scanlinesCopy (scanline.d:37-55).
 * Host pointers; pitches in bytes, may be negative; gap bytes of dst are left untouched.
 * The reference's interType/interBuf arguments are implied (both stages are fused on the GPU).
 * Returns 1 on success, 0 on failure (bool in the reference). */
int gb200_scanlines_convert(int srcType, const uint8_t* src, int srcPitch,
                            int dstType, uint8_t* dst, int dstPitch, int width, int height);
/* Same, device-resident src/dst, asynchronous on `stream`. */
int gb200_scanlines_convert_device(int srcType, const uint8_t* src, long long srcPitch,
                                   int dstType, uint8_t* dst, long long dstPitch,
                                   int width, int height, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAMUT_B200_H */
