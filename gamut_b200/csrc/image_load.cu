// image_load.cu -- Image.loadFromMemory as ONE call: decode on the GPU, convert on the GPU straight into the layout the
// caller asked for, one copy back.
//
// The reference loads in two host passes: the plugin decodes into a gapless malloc'd buffer of the file's own type, then
// Image.convertTo(applyLoadFlags(type, flags), flags as LayoutConstraints) allocates a second buffer and converts row
// by row (plugins/png.d:161-162, jpeg.d:103, qoi.d:139, qoix.d:145 -- the PERF note at plugins/qoix.d:134 asks for
// exactly this fusion). Here the decoded pixels stay in HBM, the PixelType converter (convert.cu) writes them into a
// device image of the FINAL geometry -- pitch, alignment, border, trailing pixels, vertical flip, as
// allocatePixelStorage lays it out (internals/types.d:355-540) -- and that image travels to the host once.
// SURVEY 8(f2). The decision logic restated here (host side, no pixel work):
//   identifyFormatFromMemory image.d:1037-1061      loadPNG/JPEG/QOI/QOIX flag logic  plugins/*.d
//   applyLoadFlags / computeRequestedImageComponents / validLoadFlags  internals/types.d:563-661
//   convertTo image.d:1180-1332   getAdHocLayoutConstraints image.d:1809-1905   allocatePixelStorage internals/types.d:355-540
#include "common.h"
#include "batch.h"
#include <cmath>

namespace gb {
gb200_batch* png_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, int want16, cudaStream_t st);
gb200_batch* jpeg_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int req_comps, cudaStream_t st);
gb200_batch* tga_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev, cudaStream_t st);
gb200_batch* bmp_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                              int req_comp, cudaStream_t st);
gb200_batch* qoix_decode_batch(int n, const uint8_t* const* files, const size_t* lens, const uint8_t* const* files_dev,
                               int flags, cudaStream_t st);
gb200_batch* qoi_decode_batch1(const uint8_t* data, int size, int channels, int* file_channels, cudaStream_t st);
}

namespace {

// internals/errors.d
const char* const kDecodingFailed = "Image decoding failed";
const char* const kUnidentified = "Unidentified image format";
const char* const kNoLoadSupport = "Cannot decode this image format in this build";
const char* const kTooLarge = "Can't have an image that exceeds Gamut size limitations";
const char* const kWrongComponents = "Invalid number of component for image";
const char* const kInvalidFlags = "Invalid image decoding flags";
const char* const kOutOfMemory = "Out of memory";
const char* const kUnsupportedConversion = "Unsupported image pixel type conversion";

enum { LOAD_GREYSCALE = 0x10000, LOAD_ALPHA = 0x20000, LOAD_NO_ALPHA = 0x40000, LOAD_RGB = 0x80000, LOAD_8BIT = 0x100000,
       LOAD_16BIT = 0x200000, LOAD_FP32 = 0x400000, LOAD_PREMUL = 0x1000000, LOAD_NO_PREMUL = 0x2000000 };
enum { LAYOUT_VERT_FLIPPED = 512, LAYOUT_VERT_STRAIGHT = 1024, LAYOUT_GAPLESS = 2048, LAYOUT_BORDER_MASK = 384 };

bool valid_flags(int f)                         // internals/types.d:563-578
{
    if ((f & LOAD_GREYSCALE) && (f & LOAD_RGB)) return false;
    if ((f & LOAD_ALPHA) && (f & LOAD_NO_ALPHA)) return false;
    if ((f & LOAD_PREMUL) && (f & LOAD_NO_PREMUL)) return false;
    int n = 0; if (f & LOAD_8BIT) ++n; if (f & LOAD_16BIT) ++n; if (f & LOAD_FP32) ++n;
    return n <= 1;
}
int requested_components(int f)                 // internals/types.d:588-611
{
    if (!valid_flags(f)) return 0;
    if (f & LOAD_GREYSCALE) { if (f & LOAD_ALPHA) return 2; if (f & LOAD_NO_ALPHA) return 1; }
    else if (f & LOAD_RGB) { if (f & LOAD_ALPHA) return 4; if (f & LOAD_NO_ALPHA) return 3; }
    return -1;
}
// PixelType = 3 * model + depth; models l, la, lap, rgb, rgba, rgbap (types.d:32-59)
int map_model(int t, const int (&tab)[6]) { return t < 0 ? -1 : tab[t / 3] * 3 + t % 3; }
int apply_load_flags(int t, int f)              // internals/types.d:627-661; tables types.d:351-602
{
    static const int grey[6] = {0, 1, 2, 0, 1, 2}, rgb[6] = {3, 4, 5, 3, 4, 5}, adda[6] = {1, 1, 2, 4, 4, 5};
    static const int dropa[6] = {0, 0, 0, 3, 3, 3}, premul[6] = {0, 2, 2, 3, 5, 5}, nopremul[6] = {0, 1, 1, 3, 4, 4};
    if (!valid_flags(f)) return -1;
    if (f & LOAD_GREYSCALE) t = map_model(t, grey);
    if (f & LOAD_RGB) t = map_model(t, rgb);
    if (f & LOAD_ALPHA) t = map_model(t, adda);
    if (f & LOAD_NO_ALPHA) t = map_model(t, dropa);
    if (f & LOAD_8BIT) t = t - t % 3;
    if (f & LOAD_16BIT) t = t - t % 3 + 1;
    if (f & LOAD_FP32) t = t - t % 3 + 2;
    if (f & LOAD_PREMUL) t = map_model(t, premul);
    if (f & LOAD_NO_PREMUL) t = map_model(t, nopremul);
    return t;
}
// layout accessors, internals/types.d:163-235 (layoutScanlineAlignment masks with 0x0f -- which includes the low border
// bit -- in the reference; restated, not fixed)
int lay_mult(int c) { return 1 << (c & 3); }
int lay_trailing(int c) { return (1 << ((c & 0x0C) >> 2)) - 1; }
int lay_align(int c) { return 1 << ((c >> 4) & 0x0f); }
int lay_border(int c) { return (c >> 7) & 3; }
bool lay_valid(int c)                           // internals/types.d:262-283
{
    if ((c & LAYOUT_VERT_FLIPPED) && (c & LAYOUT_VERT_STRAIGHT)) return false;
    if (c & LAYOUT_GAPLESS) if (lay_mult(c) > 1 || lay_trailing(c) > 0 || lay_align(c) > 1 || lay_border(c) > 0) return false;
    return true;
}
bool lay_compatible(int newer, int older)       // internals/types.d:236-259
{
    if ((newer & LAYOUT_GAPLESS) && !(older & LAYOUT_GAPLESS)) return false;
    if ((newer & LAYOUT_VERT_FLIPPED) && !(older & LAYOUT_VERT_FLIPPED)) return false;
    if ((newer & LAYOUT_VERT_STRAIGHT) && !(older & LAYOUT_VERT_STRAIGHT)) return false;
    return lay_mult(newer) <= lay_mult(older) && lay_trailing(newer) <= lay_trailing(older) &&
           lay_align(newer) <= lay_align(older) && lay_border(newer) <= lay_border(older);
}
int ptr_alignment_flag(size_t p)                // getPointerAlignment, internals/types.d:201-211
{
    if ((p & 127) == 0) return 112; if ((p & 63) == 0) return 96; if ((p & 31) == 0) return 80; if ((p & 15) == 0) return 64;
    if ((p & 7) == 0) return 48; if ((p & 3) == 0) return 32; if ((p & 1) == 0) return 16; return 0;
}
// getAdHocLayoutConstraints (image.d:1809-1905) of a freshly decoded image: gapless, positive pitch, one layer, no
// constraints of its own; `ptr` is the address of its first scanline.
int adhoc_of_decoded(int width, int pitch, int px, size_t ptr)
{
    (void)px;
    int c = 0;
    const int wd = width % 8 == 0 ? 8 : width % 4 == 0 ? 4 : width % 2 == 0 ? 2 : 1;      // excess pixels are 0: only the width speaks
    c |= wd == 8 ? 3 : wd == 4 ? 2 : wd == 2 ? 1 : 0;
    const int pa = ptr_alignment_flag(ptr), qa = ptr_alignment_flag((size_t)pitch);
    c |= pa < qa ? pa : qa;
    if (pitch >= 0) c |= LAYOUT_VERT_STRAIGHT;
    if (pitch <= 0) c |= LAYOUT_VERT_FLIPPED;
    c |= LAYOUT_GAPLESS;                        // pitch == |pitch| (image.d:1886)
    return c;
}

} // namespace

/* see include/gamut_b200.h */
GB_API int gb200_image_load(const uint8_t* data, size_t len, int flags, gb200_image* out)
{
    gb::clear_error();
    if (!out) return 0;
    memset(out, 0, sizeof(*out));
    out->type = -1; out->pixelAspectRatio = -1; out->resolutionY = -1;
    auto fail = [&](const char* msg) { out->error = msg; gb::set_error("%s", msg); return 0; };
    // ---- identifyFormatFromMemory (image.d:1037-1061) + loadFromStreamInternal (:1751-1772): an unknown format and a
    // format without a loader are different errors
    const int fmt = gb200_identify_format(data, len);
    if (fmt < 0) return fail(kUnidentified);
    if (fmt != GB200_FORMAT_JPEG && fmt != GB200_FORMAT_PNG && fmt != GB200_FORMAT_QOI && fmt != GB200_FORMAT_QOIX && fmt != GB200_FORMAT_BMP &&
        fmt != GB200_FORMAT_TGA)
        return fail(kNoLoadSupport);            // DDS / GIF / JXL / SQZ: detected, no decoder in this build
    if (len > 0x7fffffffu) return fail(kDecodingFailed);
    int req = requested_components(flags);
    if (req == 0 && fmt != GB200_FORMAT_TGA) return fail(kInvalidFlags);      // loadTGA does not look at the component flags itself
    if (!gb::ensure_device()) { out->error = gb200_last_error(); return 0; }      // there is no CPU fallback
    cudaStream_t st = gb::thread_stream();
    const uint8_t* f[1] = {data}; size_t l[1] = {len};
    gb200_batch* B = nullptr;
    int type = -1; float par = -1, resY = -1;
    // ---- the plugins' flag logic and the type each one adopts its buffer as
    if (fmt == GB200_FORMAT_PNG) {              // plugins/png.d:44-163
        if (req == -1) req = 0;
        bool to16 = gb200_png_is16(data, len) != 0;
        if (flags & LOAD_8BIT) to16 = false;
        if (flags & LOAD_16BIT) to16 = true;
        B = gb::png_decode_batch(1, f, l, nullptr, req, to16 ? 1 : 0, st);
        if (!B || !B->images[0].status) { delete B; return fail(kDecodingFailed); }
        const gb200_image_desc& D = B->images[0];
        const int comps = req ? req : D.file_channels;
        static const int t8[5] = {-1, GB200_l8, GB200_la8, GB200_rgb8, GB200_rgba8}, t16[5] = {-1, GB200_l16, GB200_la16, GB200_rgb16, GB200_rgba16};
        type = to16 ? t16[comps] : t8[comps];
        par = D.pixelAspectRatio == -1 ? -1.0f : D.pixelAspectRatio;
        resY = D.ppmY == -1 ? -1.0f : D.ppmY / 39.37007874f;            // convertInchesToMeters, types.d:127
    } else if (fmt == GB200_FORMAT_JPEG) {      // plugins/jpeg.d:42-104
        if (req == 2) req = -1;
        B = gb::jpeg_decode_batch(1, f, l, nullptr, req, st);
        if (!B || !B->images[0].status) {
            delete B;
            return fail(gb200_jpeg_probe(data, len) > 0 ? kNoLoadSupport : kDecodingFailed);   // non-interleaved multi-scan sequential: a valid file this path does not decode
        }
        const gb200_image_desc& D = B->images[0];
        if (D.file_channels != 1 && D.file_channels != 3 && D.file_channels != 4) { delete B; return fail(kWrongComponents); }
        const int comps = req == -1 ? D.file_channels : req;
        type = comps == 1 ? GB200_l8 : comps == 3 ? GB200_rgb8 : GB200_rgba8;
        par = D.pixelAspectRatio == -1 ? -1.0f : D.pixelAspectRatio;
        resY = D.ppmY == -1 ? -1.0f : D.ppmY;
    } else if (fmt == GB200_FORMAT_QOI) {       // plugins/qoi.d:48-140
        if (req == -1 || req == 1 || req == 2) req = 0;
        int fch = 0;
        B = gb::qoi_decode_batch1(data, (int)len, req, &fch, st);
        if (!B || !B->images[0].status) { delete B; return fail(kDecodingFailed); }
        type = (req ? req : fch) == 3 ? GB200_rgb8 : GB200_rgba8;
    } else if (fmt == GB200_FORMAT_BMP) {       // plugins/bmp.d:93-163
        if (req == -1) req = 0;
        B = gb::bmp_decode_batch(1, f, l, nullptr, req, st);
        if (!B || !B->images[0].status) { delete B; return fail(kDecodingFailed); }
        const gb200_image_desc& D = B->images[0];
        static const int t8[5] = {-1, GB200_l8, GB200_la8, GB200_rgb8, GB200_rgba8};
        type = t8[req ? req : D.file_channels];
        par = D.pixelAspectRatio == -1 ? -1.0f : D.pixelAspectRatio;
        resY = D.ppmY == -1 ? -1.0f : D.ppmY / 39.37007874f;            // convertInchesToMeters(ppmY), bmp.d:134
    } else if (fmt == GB200_FORMAT_TGA) {       // plugins/tga.d:45-105: no flag logic of its own, unknown resolution
        B = gb::tga_decode_batch(1, f, l, nullptr, st);
        if (!B || !B->images[0].status) { delete B; return fail(kDecodingFailed); }
        type = B->images[0].pixel_type;
    } else {                                    // plugins/qoix.d:64-146
        B = gb::qoix_decode_batch(1, f, l, nullptr, flags, st);
        if (!B || !B->images[0].status) { delete B; return fail(kDecodingFailed); }
        const gb200_image_desc& D = B->images[0];
        type = D.pixel_type; par = D.pixelAspectRatio; resY = D.ppmY;
    }
    const gb200_image_desc D = B->images[0];
    const int W = D.width, H = D.height;
    if (W < 0 || H < 0 || W > 16777216 || H > 16777216) { delete B; return fail(kTooLarge); }    // imageIsValidSize
    const int spx = gb200_pixel_type_size(type);
    // ---- image.convertTo(applyLoadFlags(type, flags), cast(LayoutConstraints) flags)   (image.d:1180-1332)
    const int target = apply_load_flags(type, flags);
    const int layout = flags & 0xFFFF;
    out->width = W; out->height = H; out->pixelAspectRatio = par; out->resolutionY = resY;
    if (target < 0) {
        // the reference keeps the decoded image and reports the conversion error on it
        delete B;
        return fail(kUnsupportedConversion);
    }
    if (!lay_valid(layout)) { delete B; return fail(kInvalidFlags); }      // an assert in the reference
    const size_t src_bytes = (size_t)D.pitch * H;
    // the decoded buffer as the reference would hold it: malloc'd (16-byte aligned on this ABI), gapless
    uint8_t* host_plain = nullptr;
    const bool same_type = target == type;
    bool keep = false;
    if (same_type || W == 0 || H == 0) {
        host_plain = (uint8_t*)malloc(src_bytes ? src_bytes : 1);
        if (!host_plain) { delete B; return fail(kOutOfMemory); }
        keep = lay_compatible(layout, adhoc_of_decoded(W, D.pitch, spx, (size_t)host_plain));
    }
    bool ok = true;
    if (keep) {
        // same type, compatible layout: the decoded buffer IS the image (image.d:1204-1215)
        ok = gb::cuda_ok(cudaMemcpyAsync(host_plain, D.pixels, src_bytes, cudaMemcpyDeviceToHost, st), "pixels to host", __FILE__, __LINE__) &&
             gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);
        delete B;
        if (!ok) { free(host_plain); out->error = gb200_last_error(); return 0; }
        out->alloc = host_plain; out->alloc_bytes = src_bytes ? src_bytes : 1; out->data = host_plain; out->type = type; out->pitch = D.pitch; out->layout = layout;
        return 1;
    }
    free(host_plain);
    // ---- allocatePixelStorage (internals/types.d:355-540), one layer
    const int border = lay_border(layout), align = lay_align(layout), trailing = lay_trailing(layout), mult = lay_mult(layout);
    const int right_pad = (W + border + mult - 1) / mult * mult - (W + border);
    int border_right = border + right_pad; if (border_right < trailing) border_right = trailing;
    const long long actual_w = (long long)border + W + border_right, actual_h = (long long)border + H + border;
    const int dpx = gb200_pixel_type_size(target);
    long long pitch = ((long long)dpx * actual_w + align - 1) / align * align;
    const long long bonus = same_type ? 0 : (long long)W * gb200_pixel_type_size(gb200_scanlines_inter_type(type, target));   // image.d:1233-1236
    const long long need = pitch * actual_h + (align - 1) + bonus;
    if (need > 0x7FFFFFFFLL) { delete B; return fail(kOutOfMemory); }
    uint8_t* area = (uint8_t*)malloc(need ? (size_t)need : 1);
    if (!area) { delete B; return fail(kOutOfMemory); }
    size_t first = (size_t)area + (size_t)bonus + (size_t)pitch * border + (size_t)dpx * border;
    first = (first + align - 1) / align * align;
    const size_t first_off = first - (size_t)area;               // offset of the first STORED row (top row if not flipped)
    // the rows travel as one block [first_off, first_off + pitch * H); on the device that block starts at offset 0
    const size_t block = (size_t)pitch * H;
    long long final_pitch = pitch; size_t data_off = first_off;
    if ((layout & LAYOUT_VERT_FLIPPED) && pitch > 0) {           // applyVFlipConstraintsToScanlinePointers :303-320
        if (H >= 2) data_off += (size_t)pitch * (H - 1);
        final_pitch = -pitch;
    }
    if (block) {
        gb::DevBuf d_img(block);
        if (!d_img.p) { free(area); delete B; return fail(kOutOfMemory); }
        uint8_t* d_first = d_img.as<uint8_t>() + (data_off - first_off);      // device address of scanline 0
        ok = gb::dev_fill_async(d_img.p, 0, block, st);
        // scanlinesConvert / scanlinesCopy (image.d:1262-1300) on the device, straight into the final geometry
        ok = ok && gb200_scanlines_convert_device(type, D.pixels, D.pitch, target, d_first, final_pitch, W, H, st);
        ok = ok && gb::cuda_ok(cudaMemcpyAsync(area + first_off, d_img.p, block, cudaMemcpyDeviceToHost, st), "image to host", __FILE__, __LINE__);
        ok = ok && gb::cuda_ok(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__);   // before d_img returns to the pool
        if (!ok) cudaStreamSynchronize(st);
    }
    delete B;
    if (!ok) { free(area); out->error = kUnsupportedConversion; return 0; }
    out->alloc = area; out->alloc_bytes = need ? (size_t)need : 1; out->data = area + data_off; out->type = target; out->pitch = (int)final_pitch; out->layout = layout;
    return 1;
}
