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
enum gb200_image_format { GB200_FORMAT_JPEG = 0, GB200_FORMAT_PNG = 1, GB200_FORMAT_QOI = 2, GB200_FORMAT_QOIX = 3,    /* ImageFormat, types.d:14-28 */
                          GB200_FORMAT_DDS = 4, GB200_FORMAT_TGA = 5, GB200_FORMAT_GIF = 6, GB200_FORMAT_BMP = 7, GB200_FORMAT_JXL = 8, GB200_FORMAT_SQZ = 9 };

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
int         gb200_copy_to_host(void* dst_host, const void* src_dev, size_t bytes);    /* synchronous */
int         gb200_copy_to_device(void* dst_dev, const void* src_host, size_t bytes);  /* synchronous */
/* Download by the SMs instead of a copy engine: 16-byte stores straight into PINNED host memory (gb200_host_alloc),
 * asynchronous on `stream`; pointers and size multiples of 16. */
int         gb200_download_by_kernel(void* dst_pinned, const void* src_dev, size_t bytes, void* stream);

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

/* ---- batched, device-resident decode (one call decodes n independent images) ----
 * The result object owns the device memory of every decoded image; descriptors are valid until
 * gb200_batch_free(). A failed image has status 0 and pixels NULL -- it never poisons its neighbours
 * (the reference's per-Image sticky error, image.d:1563). The call returns after the work on `stream`
 * has completed (statuses and sizes are read back). */
typedef struct gb200_batch gb200_batch;
typedef struct gb200_image_desc {
    uint8_t* pixels;        /* device pointer, gapless rows */
    int width, height;
    int channels;           /* channels of the decoded buffer */
    int file_channels;      /* channels reported by the file (stb `comp`, jpgd `actual_comps`) */
    int bits;               /* bits per channel: 8 or 16 */
    int pixel_type;         /* gb200_pixel_type of the decoded buffer */
    int pitch;              /* bytes per row */
    int status;             /* 1 decoded, 0 failed */
    float ppmX, ppmY, pixelAspectRatio;   /* PNG pHYs (-1 unknown); JPEG: ppmY holds dpiY */
} gb200_image_desc;

int                     gb200_batch_count(const gb200_batch* b);
const gb200_image_desc* gb200_batch_images(const gb200_batch* b);
void                    gb200_batch_free(gb200_batch* b);
/* Copies every decoded image to host memory: image i lands at dst_host + i*stride (gapless rows). One
 * asynchronous copy per image on the batch's stream and one synchronisation; dst_host should come from
 * gb200_host_alloc (pinned) for PCIe-speed copies. Failed images are skipped. Returns 1 / 0. */
int                     gb200_batch_download(const gb200_batch* b, uint8_t* dst_host, size_t stride);
/* Device time of each phase of the call (CUDA events on the call's stream), ms. PNG: [0] IDAT gather /
 * H2D, [1] inflate, [2] unfilter, [3] finish. JPEG: [0] upload, [1] Huffman, [2] IDCT+colour. QOIX: [0] upload,
 * [1] LZ4, [2] opcode decode. */
void                    gb200_batch_timing(const gb200_batch* b, float* phase_ms8, double* host_parse_ms);

/* Host-to-host batched decode with the transfers overlapped: files[i]/lens[i] in host memory, image i delivered to
 * dst_host + i*dst_stride (gapless rows; dst_host should be pinned -- gb200_host_alloc -- or the copies are staged by
 * the driver). The batch is cut into sub-batches (sub_batch images each, 0 = automatic); the pixels of one sub-batch
 * travel back while the next one is uploaded and decoded. format: GB200_FORMAT_JPEG (arg = req_comps, -1 keep),
 * GB200_FORMAT_PNG (arg = req_comp, want16 as in gb200_png_decode_batch), GB200_FORMAT_QOIX (arg = LoadFlags),
 * GB200_FORMAT_BMP (arg = req_comp).
 * descs[i] is filled like gb200_batch_images() except that `pixels` is the HOST address of the image (NULL = failed; an
 * image larger than dst_stride fails). Synchronous, thread-safe. This is the batch form of the reference's codec
 * calls -- bytes in, pixels out (plugins/png.d:108, jpeg.d:62, qoix.d:116). Returns 1 / 0. */
int gb200_decode_batch_host(int format, int n, const uint8_t* const* files, const size_t* lens, int arg, int want16,
                            uint8_t* dst_host, size_t dst_stride, gb200_image_desc* descs, int sub_batch);

/* ---- PNG: source/gamut/codecs/stbdec.d (stb_image PNG path) + miniz inflate ---- */
/* stbi__png_is16 (stbdec.d:2090-2110): 1 if the file stores 16-bit samples. Host-only header scan. */
int gb200_png_is16(const uint8_t* data, size_t len);
/* stbi_load_from_callbacks (want16 = 0, stbdec.d:725) / stbi_load_16_from_callbacks (want16 = 1,
 * stbdec.d:713) over a memory buffer. req_comp 0 = keep the file's channels, 1..4 = force.
 * Returns malloc()'d host pixels (free with gb200_free) or NULL. *comp = channels in the file. */
uint8_t* gb200_png_load(const uint8_t* data, size_t len, int req_comp, int want16,
                        int* width, int* height, int* comp, float* ppmX, float* ppmY, float* pixelRatio);
/* Batched PNG decode: files[i]/lens[i] are HOST copies of the files (chunk headers are walked on the
 * host); files_dev, if not NULL, holds device-resident copies of the same bytes (no H2D copy is made
 * then; every device copy must be followed by >= 16 readable bytes -- the kernels fetch whole 16-byte vectors;
 * the same holds for the JPEG and QOIX batch entry points). want16: 0 / 1 as above, -1 = auto per file (what
 * loadPNG does, plugins/png.d:77-79). */
gb200_batch* gb200_png_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                    const uint8_t* const* files_dev, int req_comp, int want16, void* stream);
/* Kernel-level entry for the row unfilter alone (stbi__create_png_image_raw, stbdec.d:1406-1547,
 * 8/16-bit samples, img_n == out_n): `raw` = n_images inflated streams of (1+row_bytes)*height bytes,
 * device-resident, `raw_stride` apart; out = n_images * row_bytes * height. bpp = channels*bytes.
 * status_dev (device int per image, may be NULL) is set to 0 for an image with a filter byte > 4.
 * For the fast path `raw` and `raw_stride` should be multiples of 16 with >= 16 readable bytes after each
 * stream (otherwise a slower generic kernel runs). */
int gb200_png_unfilter_device(const uint8_t* raw, size_t raw_stride, uint8_t* out, size_t out_stride,
                              int n_images, int row_bytes, int height, int bpp, int* status_dev, void* stream);
/* Kernel-level entry for inflate alone: n zlib (parse_header=1) or raw deflate streams, device-resident,
 * each 4-byte aligned with >= 16 readable bytes after its end. out_lens/statuses are device arrays
 * (status 0 ok, 1 output buffer too small, 2 corrupt). */
int gb200_inflate_device(int n, const uint8_t* const* in_dev, const uint32_t* in_lens,
                         uint8_t* const* out_dev, const uint32_t* out_caps, int parse_header,
                         uint32_t* out_lens_dev, int* statuses_dev, void* stream);
/* Inflate engine selection for tests and profiling: 1 = block-parallel pipeline with the one-warp-per-stream
 * decoder for whatever it does not accept (default), 0 = one warp per stream only. Results are identical. */
void gb200_inflate_set_mode(int parallel);

/* ---- JPEG: source/gamut/codecs/jpegload.d (jpgd port), baseline / extended-sequential and progressive Huffman ---- */
/* decompress_jpeg_image_from_stream (jpegload.d:3720-3808) over a memory buffer. req_comps: -1 keep,
 * 1, 3 or 4. Returns malloc()'d host pixels or NULL. Sequential files with one interleaved scan and progressive (SOF2)
 * files are decoded; non-interleaved multi-scan sequential files fail (see gb200_jpeg_probe). pixelAspectRatio / dotsPerInchY are NaN when the file carries no JFIF/EXIF density (the
 * reference's D float members are never assigned in that case). */
uint8_t* gb200_jpeg_load(const uint8_t* data, size_t len, int req_comps, int* width, int* height,
                         int* actual_comps, float* pixelAspectRatio, float* dotsPerInchY);
gb200_batch* gb200_jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                     const uint8_t* const* files_dev, int req_comps, void* stream);
/* ---- BMP encode: saveBMP (plugins/bmp.d:166-194) -> write_bmp (codecs/bmpenc.d:25-113) ----
 * type = gb200_pixel_type of the rows: rgb8 (24-bit file) or rgba8 (32-bit BI_BITFIELDS file, V4 header); sides
 * 1..32767; pitchBytes signed, `pixels` = the first scanline; ppmX / ppmY = Image.pixelsPerMeterX / Y (-1 = unknown). The
 * file equals the one write_bmp writes except for the row padding of 24-bit files, which the reference takes from an
 * uninitialised buffer (bmpenc.d:40-43) and this writer sets to zero. */
typedef struct gb200_bmp_desc { int32_t width, height, pitchBytes, type; float ppmX, ppmY; } gb200_bmp_desc;
uint8_t* gb200_bmp_encode(const uint8_t* pixels, const gb200_bmp_desc* desc, int* out_len);
size_t gb200_bmp_encode_size(const gb200_bmp_desc* desc);
/* ---- TGA ----
 * TGADecoder.getImageInfo + decodeImage (codecs/tga.d:313-588) as loadTGA calls them (plugins/tga.d:45-105): grey,
 * grey + alpha, 15/16-bit and 24/32-bit colour, colour-mapped files (8/16-bit indices; 8/15/16/24/32-bit entries), raw
 * or run-length coded, bottom-up or top-down. Returns malloc'd pixels (gb200_free) or NULL; *comp = components of the
 * decoded image (1 = l8, 2 = la8, 3 = rgb8, 4 = rgba8; the decoder has no req_comp, the plugin converts afterwards). */
uint8_t* gb200_tga_load(const uint8_t* data, size_t len, int* width, int* height, int* comp);
/* Batched, device-resident output. files[i] (host bytes) are always needed -- header and palette are read on the host;
 * files_dev, when not NULL, holds the same bytes on the device and saves the upload. */
gb200_batch* gb200_tga_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                    const uint8_t* const* files_dev, void* stream);
/* ---- TGA encode: saveTGA (plugins/tga.d:123-149) -> TGAEncoder (codecs/tga.d:62-292), run-length coding on ----
 * type = gb200_pixel_type of the rows: l8 / la8 / rgb8 / rgba8 (written as a 24- or 32-bit file, bottom row first);
 * pitchBytes signed, `pixels` = the first scanline. The file is byte-identical to the one saveTGA writes. */
typedef struct gb200_tga_desc { int32_t width, height, pitchBytes, type; } gb200_tga_desc;
uint8_t* gb200_tga_encode(const uint8_t* pixels, const gb200_tga_desc* desc, int* out_len);
size_t gb200_tga_encode_bound(const gb200_tga_desc* desc);
int gb200_tga_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_tga_desc* descs, uint8_t* const* out_dev,
                                  int* out_len, void* stream);
/* ---- BMP ----
 * stbi_load_from_callbacks on a BMP file (codecs/stbdec.d:725 -> stbi__bmp_load :2263-2466), as loadBMP calls it
 * (plugins/bmp.d:112): 1/4/8-bit palettes, 16/32-bit bit fields, 24/32-bit BGR(A), bottom-up and top-down, OS/2 and
 * V3/V4/V5 headers; req_comp 0 (as stored: 3 or 4) or 1..4. Returns malloc'd pixels (gb200_free) or NULL; *comp = the
 * file's own channel count; ppm values are the header's pixels per metre (-1 unknown). */
uint8_t* gb200_bmp_load(const uint8_t* data, size_t len, int req_comp, int* width, int* height, int* comp,
                        float* ppmX, float* ppmY, float* pixelRatio);
gb200_batch* gb200_bmp_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                    const uint8_t* const* files_dev, int req_comp, void* stream);
/* Image.identifyFormatFromMemory (image.d:1037-1061): the detect procs of all ten plugins in ImageFormat order, TGA
 * last. Returns a gb200_image_format value or -1 (unknown). Host only. */
int gb200_identify_format(const uint8_t* data, size_t len);

/* Classifies a file without decoding it (host-only marker walk): 0 = decodable by this path (baseline / extended
 * sequential Huffman with one interleaved scan, or progressive SOF2 -- init_progressive, jpegload.d:3299-3683), 2 =
 * sequential but non-interleaved multi-scan, -1 = not a valid JPEG. gb200_jpeg_load sets gb200_last_error() to a message
 * starting with "unsupported:" for 2, so that callers can route those files to another decoder. (1 used to mean
 * "progressive, unsupported" and is no longer returned.) */
int gb200_jpeg_probe(const uint8_t* data, size_t len);

/* ---- QOI: source/gamut/codecs/qoi.d ---- */
typedef struct gb200_qoi_desc { uint32_t width, height; uint8_t channels, colorspace; } gb200_qoi_desc;  /* qoi.d:215-222 minus pitchBytes */
/* qoi_decode (qoi.d:448-550). channels: 0 = as stored, 3 or 4. malloc()'d host pixels or NULL. */
uint8_t* gb200_qoi_decode(const uint8_t* data, int size, gb200_qoi_desc* desc, int channels);
/* ---- QOI encode: qoi_encode (qoi.d:295-426) as saveQOI calls it (plugins/qoi.d:150-185) ----
 * `pixels` = the first scanline of an rgb8 (channels 3) or rgba8 (channels 4) image on the host, pitchBytes signed (the
 * reference's qoi_desc.pitchBytes, qoi.d:220; negative for a vertically flipped Image). The stream is byte-identical to
 * qoi_encode's. Returns malloc()'d bytes (gb200_free) and *out_len, or NULL where the reference returns null. */
uint8_t* gb200_qoi_encode(const uint8_t* pixels, const gb200_qoi_desc* desc, int pitchBytes, int* out_len);
/* Upper bound of the stream length for a device output buffer (qoi.d:313-315, rounded up). */
size_t gb200_qoi_encode_bound(const gb200_qoi_desc* desc);
/* Batched, device-resident: pixels_dev[i] (first scanline, pitches[i] signed) -> out_dev[i] (16-byte aligned, >=
 * gb200_qoi_encode_bound bytes); out_len[i] = stream length, 0 for an image the encoder refuses. Returns 1 when the
 * batch ran. */
int gb200_qoi_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_qoi_desc* descs, const int* pitches,
                                  uint8_t* const* out_dev, int* out_len, void* stream);

/* ---- QOIX (+LZ4): source/gamut/plugins/qoix.d, codecs/{qoi2avg,qoiplane,qoiplane10,qoi10b,lz4}.d ---- */
typedef struct gb200_qoix_desc {        /* qoi_desc, qoi2avg.d:276-287 */
    uint32_t width, height;
    int32_t  pitchBytes;
    uint8_t  channels, bitdepth, colorspace, compression;
    float    pixelAspectRatio, resolutionY;
} gb200_qoix_desc;
/* qoix_lz4_decode (plugins/qoix.d:350-473): container + optional LZ4 + sub-codec dispatch. `flags` are
 * LoadFlags (only validated, as in the reference); *decodedType receives the stream's own PixelType.
 * All four sub-codecs are built: QOI-Plane10 (10-bit L/LA), QOI-10b (10-bit RGB/RGBA), QOI-Plane (8-bit L/LA) and
 * QOI2AVG (8-bit RGB/RGBA), each with and without the LZ4 wrapper. */
uint8_t* gb200_qoix_decode(const uint8_t* data, int size, gb200_qoix_desc* desc, int flags, int* decodedType);
gb200_batch* gb200_qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens,
                                     const uint8_t* const* files_dev, int flags, void* stream);

/* ---- QOIX encode: source/gamut/plugins/qoix.d:251-339, codecs/qoiplane10.d:99-314, qoiplane.d:109-375,
 * qoi2avg.d:376-617 ----
 * qoix_lz4_encode for the images it hands to qoiplane10_encode (bitdepth 10, 1 or 2 channels, 16-bit samples), to
 * qoiplane_encode (bitdepth 8, 1 or 2 channels), to qoix_encode / QOI2AVG (bitdepth 8, 3 or 4 channels) and to
 * qoi10b_encode / QOI-10b (bitdepth 10, 3 or 4 channels, 16-bit samples); desc->
 * compression must be 0, pitchBytes is honoured (not negative). The stream is bit-identical to the reference
 * sub-encoder's; the LZ4 stage (LZ4_compress, kept by the reference only when it makes the file smaller) is not built,
 * so the result always has compression = 0 -- a valid QOIX file that every decoder of the format reads. bitdepth 10
 * with 3 or 4 channels (16-bit samples) goes to qoi10b_encode / QOI-10b (qoi10b.d:136-500): all four sub-encoders of
 * qoix_lz4_encode are built. Returns malloc()'d bytes (gb200_free) or NULL. */
uint8_t* gb200_qoix_encode(const uint8_t* pixels, const gb200_qoix_desc* desc, int* out_len);
/* Upper bound of the stream length for a device output buffer (qoiplane10.d:112-116, rounded up). */
size_t gb200_qoix_encode_bound(const gb200_qoix_desc* desc);
/* Batched, device-resident: pixels_dev[i] -> out_dev[i] (16-byte aligned, >= gb200_qoix_encode_bound bytes); out_len[i] =
 * stream length, 0 for an image the encoder refuses. Returns 1 when the batch ran. */
int gb200_qoix_encode_batch_device(int n, const uint8_t* const* pixels_dev, const gb200_qoix_desc* descs,
                                   uint8_t* const* out_dev, int* out_len, void* stream);

/* ---- Image.loadFromMemory in one call (image.d:886-901 -> loadPNG/JPEG/QOI/QOIX -> convertTo) ----
 * Replaces the body of the four LoadImageProc plugins (plugin.d:30; plugins/png.d:44, jpeg.d:42, qoi.d:48, qoix.d:64)
 * INCLUDING their closing image.convertTo(applyLoadFlags(type, flags), cast(LayoutConstraints) flags): the file is
 * decoded on the GPU, converted on the GPU straight into the PixelType and LayoutConstraints that `flags` ask for
 * (pitch, scanline alignment, border, trailing pixels, vertical flip as allocatePixelStorage lays them out,
 * internals/types.d:355-540) and copied to the host once -- the fusion the reference's PERF note at plugins/qoix.d:134
 * asks for. `flags` = LoadFlags | LayoutConstraints exactly as passed to Image.loadFromMemory.
 * On success (1): alloc = malloc()'d area of alloc_bytes to adopt as Image._allocArea (free with gb200_free), data = first scanline
 * (Image._data), pitch signed (negative when flipped), type = gb200_pixel_type, layout = flags & 0xFFFF, plus the
 * resolution fields. On failure (0): error = the static string the reference would set with image.error(kStr...). */
typedef struct gb200_image {
    void*    alloc;
    size_t   alloc_bytes;
    uint8_t* data;
    int      width, height;
    int      type;
    int      pitch;
    int      layout;
    float    pixelAspectRatio, resolutionY;
    const char* error;
} gb200_image;
int gb200_image_load(const uint8_t* data, size_t len, int flags, gb200_image* out);

#ifdef __cplusplus
}
#endif
#endif /* GAMUT_B200_H */
