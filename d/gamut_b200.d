/**
D binding of the gamut_b200 C ABI (include/gamut_b200.h) and drop-in replacements for the four load
procs of the reference's plugin table (source/gamut/plugin.d:30-53, 111-123).

betterC compatible, `nothrow @nogc`. NOT COMPILED in the build image (no dmd/ldc2/gdc there): delivered
as source for a maintainer; every call below mirrors a ctypes call that *is* exercised by the test-suite
(gamut_b200/codecs.py, gamut_b200/image.py), and the epilogues are the reference's own field adoption
(plugins/png.d:108-162, jpeg.d:83-103, qoi.d:112-139, qoix.d:122-145) kept verbatim in meaning.

Usage (see INTEGRATION.md): add this file to the dub package, link libgamut_b200.so, and in
source/gamut/plugin.d point `loadProc` of the PNG/JPEG/QOI/QOIX plugins at the functions below.
*/
module gamut.gamut_b200;

nothrow @nogc:

import core.stdc.stdlib : malloc, free;
import core.stdc.stdio : SEEK_END;

import gamut.types;
import gamut.io;
import gamut.image;
import gamut.internals.errors;
import gamut.internals.types;

// ---------------------------------------------------------------------------------------------
// extern(C) prototypes -- one line per entry point of include/gamut_b200.h
// ---------------------------------------------------------------------------------------------
extern(C)
{
    int gb200_init();
    const(char)* gb200_version();
    const(char)* gb200_last_error();
    long gb200_launch_count();
    void gb200_free(void* p);
    void* gb200_host_alloc(size_t bytes);
    void gb200_host_free(void* p);

    int gb200_pixel_type_size(int type);
    int gb200_scanlines_inter_type(int srcType, int dstType);

    /// scanlinesConvert / scanlinesCopy (scanline.d:70-121, :37-55); host pointers, signed pitches.
    int gb200_scanlines_convert(int srcType, const(ubyte)* src, int srcPitch,
                                int dstType, ubyte* dst, int dstPitch, int width, int height);
    int gb200_scanlines_convert_device(int srcType, const(ubyte)* src, long srcPitch,
                                       int dstType, ubyte* dst, long dstPitch,
                                       int width, int height, void* stream);

    /// stbi__png_is16 (stbdec.d:2090-2110)
    int gb200_png_is16(const(ubyte)* data, size_t len);
    /// stbi_load_from_callbacks / stbi_load_16_from_callbacks (stbdec.d:713-735) over a memory buffer.
    ubyte* gb200_png_load(const(ubyte)* data, size_t len, int req_comp, int want16,
                          int* width, int* height, int* comp, float* ppmX, float* ppmY, float* pixelRatio);

    /// decompress_jpeg_image_from_stream (jpegload.d:3720-3808) over a memory buffer.
    ubyte* gb200_jpeg_load(const(ubyte)* data, size_t len, int req_comps, int* width, int* height,
                           int* actual_comps, float* pixelAspectRatio, float* dotsPerInchY);

    struct gb200_qoi_desc { uint width, height; ubyte channels, colorspace; }
    /// qoi_decode (qoi.d:448-550)
    ubyte* gb200_qoi_decode(const(ubyte)* data, int size, gb200_qoi_desc* desc, int channels);

    struct gb200_qoix_desc
    {
        uint width, height;
        int pitchBytes;
        ubyte channels, bitdepth, colorspace, compression;
        float pixelAspectRatio, resolutionY;
    }
    /// qoix_lz4_decode (plugins/qoix.d:350-473)
    ubyte* gb200_qoix_decode(const(ubyte)* data, int size, gb200_qoix_desc* desc, int flags, int* decodedType);

    // batched, device-resident decoders (no reference counterpart: the new capability)
    struct gb200_batch;
    struct gb200_image_desc
    {
        ubyte* pixels;
        int width, height, channels, file_channels, bits, pixel_type, pitch, status;
        float ppmX, ppmY, pixelAspectRatio;
    }
    int gb200_batch_count(const(gb200_batch)* b);
    const(gb200_image_desc)* gb200_batch_images(const(gb200_batch)* b);
    void gb200_batch_free(gb200_batch* b);
    gb200_batch* gb200_png_decode_batch(int n, const(ubyte*)* files, const(size_t)* lens,
                                        const(ubyte*)* files_dev, int req_comp, int want16, void* stream);
    gb200_batch* gb200_jpeg_decode_batch(int n, const(ubyte*)* files, const(size_t)* lens,
                                         const(ubyte*)* files_dev, int req_comps, void* stream);
    gb200_batch* gb200_qoix_decode_batch(int n, const(ubyte*)* files, const(size_t)* lens,
                                         const(ubyte*)* files_dev, int flags, void* stream);
    /// qoix_lz4_encode (plugins/qoix.d:251) for 1/2-channel images: the qoiplane10_encode (10-bit) / qoiplane_encode
    /// (8-bit) stream, compression 0
    ubyte* gb200_qoix_encode(const(ubyte)* pixels, const(gb200_qoix_desc)* desc, int* out_len);
    size_t gb200_qoix_encode_bound(const(gb200_qoix_desc)* desc);
    int gb200_qoix_encode_batch_device(int n, const(ubyte*)* pixels_dev, const(gb200_qoix_desc)* descs,
                                       const(ubyte*)* out_dev, int* out_len, void* stream);
    /// qoi_encode (codecs/qoi.d:295) for rgb8 / rgba8 rows with a signed pitch: the reference encoder's stream
    ubyte* gb200_qoi_encode(const(ubyte)* pixels, const(gb200_qoi_desc)* desc, int pitchBytes, int* out_len);
    size_t gb200_qoi_encode_bound(const(gb200_qoi_desc)* desc);
    int gb200_qoi_encode_batch_device(int n, const(ubyte*)* pixels_dev, const(gb200_qoi_desc)* descs, const(int)* pitches,
                                      const(ubyte*)* out_dev, int* out_len, void* stream);
    /// TGADecoder.getImageInfo + decodeImage (codecs/tga.d:313-588): malloc'd l8 / la8 / rgb8 / rgba8 pixels by *comp
    ubyte* gb200_tga_load(const(ubyte)* data, size_t len, int* width, int* height, int* comp);
    gb200_batch* gb200_tga_decode_batch(int n, const(ubyte*)* files, const(size_t)* lens, const(ubyte*)* files_dev, void* stream);
    /// saveTGA -> TGAEncoder (codecs/tga.d:62-292): the run-length file, byte for byte
    struct gb200_tga_desc { int width, height, pitchBytes, type; }
    ubyte* gb200_tga_encode(const(ubyte)* pixels, const(gb200_tga_desc)* desc, int* out_len);
    size_t gb200_tga_encode_bound(const(gb200_tga_desc)* desc);
    int gb200_tga_encode_batch_device(int n, const(ubyte*)* pixels_dev, const(gb200_tga_desc)* descs, const(ubyte*)* out_dev,
                                      int* out_len, void* stream);
    /// saveBMP -> write_bmp (codecs/bmpenc.d:25-113): rgb8 / rgba8 rows -> a BMP file with a V4 header
    struct gb200_bmp_desc { int width, height, pitchBytes, type; float ppmX, ppmY; }
    ubyte* gb200_bmp_encode(const(ubyte)* pixels, const(gb200_bmp_desc)* desc, int* out_len);
    size_t gb200_bmp_encode_size(const(gb200_bmp_desc)* desc);
    int gb200_copy_to_host(void* dst_host, const(void)* src_dev, size_t bytes);
    int gb200_copy_to_device(void* dst_dev, const(void)* src_host, size_t bytes);
    int gb200_download_by_kernel(void* dst_pinned, const(void)* src_dev, size_t bytes, void* stream);
    void* gb200_device_alloc(size_t bytes);
    void gb200_device_free(void* p);
    void gb200_device_trim();
    int gb200_sm_count();
    int gb200_batch_download(const(gb200_batch)* b, ubyte* dst_host, size_t stride);
    int gb200_jpeg_probe(const(ubyte)* data, size_t len);
    /// stbi_load_from_callbacks on a BMP (stbdec.d:725 -> :2263), as loadBMP calls it (plugins/bmp.d:112)
    ubyte* gb200_bmp_load(const(ubyte)* data, size_t len, int req_comp, int* width, int* height, int* comp,
                          float* ppmX, float* ppmY, float* pixelRatio);
    gb200_batch* gb200_bmp_decode_batch(int n, const(ubyte*)* files, const(size_t)* lens,
                                        const(ubyte*)* files_dev, int req_comp, void* stream);
    /// Image.identifyFormatFromMemory (image.d:1037-1061): ImageFormat value or -1
    int gb200_identify_format(const(ubyte)* data, size_t len);
    struct gb200_image { void* alloc; size_t alloc_bytes; ubyte* data; int width, height, type, pitch, layout; float pixelAspectRatio, resolutionY; const(char)* error; }
    int gb200_image_load(const(ubyte)* data, size_t len, int flags, gb200_image* out_);
    int gb200_decode_batch_host(int format, int n, const(ubyte*)* files, const(size_t)* lens, int arg, int want16,
                                ubyte* dst_host, size_t dst_stride, gb200_image_desc* descs, int sub_batch);
    void gb200_batch_timing(const(gb200_batch)* b, float* phase_ms8, double* host_parse_ms);

    // kernel-level entries (device-resident data): row unfilter and inflate on their own
    int gb200_png_unfilter_device(const(ubyte)* raw, size_t raw_stride, ubyte* out_, size_t out_stride,
                                  int n_images, int row_bytes, int height, int bpp, int* status_dev, void* stream);
    int gb200_inflate_device(int n, const(ubyte*)* in_dev, const(uint)* in_lens, ubyte** out_dev, const(uint)* out_caps,
                             int parse_header, uint* out_lens_dev, int* statuses_dev, void* stream);
    void gb200_inflate_set_mode(int parallel);
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------

/// Reads the whole stream into a malloc'd buffer, the way loadQOI does (plugins/qoi.d:54-95).
private ubyte* slurp(IOStream* io, IOHandle handle, out int len, ref Image image) @trusted
{
    if (io.seek(handle, 0, SEEK_END) != 0) { image.error(kStrImageDecodingIOFailure); return null; }
    len = cast(int) io.tell(handle);
    if (!io.rewind(handle)) { image.error(kStrImageDecodingIOFailure); return null; }
    ubyte* buf = cast(ubyte*) malloc(len ? len : 1);
    if (buf is null) { image.error(kStrImageDecodingMallocFailure); return null; }
    if (len != io.read(buf, 1, len, handle))
    {
        free(buf);
        image.error(kStrImageDecodingIOFailure);
        return null;
    }
    return buf;
}

// ---------------------------------------------------------------------------------------------
// replacement load procs (LoadImageProc, plugin.d:30)
// ---------------------------------------------------------------------------------------------

/// Replaces loadPNG (plugins/png.d:44-163).
void loadPNG_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    bool is16bit = gb200_png_is16(buf, len) != 0;
    int requestedComp = computeRequestedImageComponents(flags);
    if (requestedComp == 0) { image.error(kStrInvalidFlags); return; }
    if (requestedComp == -1) requestedComp = 0;

    bool decodeTo16bit = is16bit;
    if (flags & LOAD_8BIT) decodeTo16bit = false;
    if (flags & LOAD_16BIT) decodeTo16bit = true;

    int width, height, components;
    float ppmX = -1, ppmY = -1, pixelRatio = -1;
    ubyte* decoded = gb200_png_load(buf, len, requestedComp, decodeTo16bit ? 1 : 0,
                                    &width, &height, &components, &ppmX, &ppmY, &pixelRatio);
    if (requestedComp != 0) components = requestedComp;
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (!imageIsValidSize(1, width, height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    image._allocArea = decoded;      // malloc'd by the C side, freed by deallocatePixelStorage
    image._width = width;
    image._height = height;
    image._data = decoded;
    image._pitch = width * components * (decodeTo16bit ? 2 : 1);
    image._pixelAspectRatio = (pixelRatio == -1) ? GAMUT_UNKNOWN_ASPECT_RATIO : pixelRatio;
    image._resolutionY = (ppmY == -1) ? GAMUT_UNKNOWN_RESOLUTION : convertInchesToMeters(ppmY);
    image._layoutConstraints = LAYOUT_DEFAULT;
    image._layerCount = 1;
    image._layerOffset = 0;
    static immutable PixelType[5] t8  = [PixelType.unknown, PixelType.l8, PixelType.la8, PixelType.rgb8, PixelType.rgba8];
    static immutable PixelType[5] t16 = [PixelType.unknown, PixelType.l16, PixelType.la16, PixelType.rgb16, PixelType.rgba16];
    image._type = decodeTo16bit ? t16[components] : t8[components];
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces loadTGA (plugins/tga.d:45-105).
void loadTGA_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    int width, height, components;
    ubyte* decoded = gb200_tga_load(buf, len, &width, &height, &components);
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (!imageIsValidSize(1, width, height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    static immutable PixelType[5] t8 = [PixelType.unknown, PixelType.l8, PixelType.la8, PixelType.rgb8, PixelType.rgba8];
    image._type = t8[components];
    image._allocArea = decoded;
    image._width = width;
    image._height = height;
    image._data = decoded;
    image._pitch = width * components;
    image._pixelAspectRatio = GAMUT_UNKNOWN_ASPECT_RATIO;
    image._resolutionY = GAMUT_UNKNOWN_RESOLUTION;
    image._layoutConstraints = LAYOUT_DEFAULT;
    image._layerCount = 1;
    image._layerOffset = 0;
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces loadBMP (plugins/bmp.d:93-163).
void loadBMP_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    int requestedComp = computeRequestedImageComponents(flags);
    if (requestedComp == 0) { image.error(kStrInvalidFlags); return; }
    if (requestedComp == -1) requestedComp = 0;

    int width, height, components;
    float ppmX = -1, ppmY = -1, pixelRatio = -1;
    ubyte* decoded = gb200_bmp_load(buf, len, requestedComp, &width, &height, &components, &ppmX, &ppmY, &pixelRatio);
    if (requestedComp != 0) components = requestedComp;
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (!imageIsValidSize(1, width, height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    image._allocArea = decoded;
    image._width = width;
    image._height = height;
    image._data = decoded;
    image._pitch = width * components;
    image._pixelAspectRatio = (pixelRatio == -1) ? GAMUT_UNKNOWN_ASPECT_RATIO : pixelRatio;
    image._resolutionY = (ppmY == -1) ? GAMUT_UNKNOWN_RESOLUTION : convertInchesToMeters(ppmY);
    image._layoutConstraints = LAYOUT_DEFAULT;
    image._layerCount = 1;
    image._layerOffset = 0;
    static immutable PixelType[5] t8 = [PixelType.unknown, PixelType.l8, PixelType.la8, PixelType.rgb8, PixelType.rgba8];
    image._type = t8[components];
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces loadJPEG (plugins/jpeg.d:42-104).
void loadJPEG_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int requestedComp = computeRequestedImageComponents(flags);
    if (requestedComp == 0) { image.error(kStrInvalidFlags); return; }
    if (requestedComp == 2) requestedComp = -1;

    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    int width, height, actualComp;
    float pixelAspectRatio, dotsPerInchY;
    ubyte* decoded = gb200_jpeg_load(buf, len, requestedComp, &width, &height, &actualComp,
                                     &pixelAspectRatio, &dotsPerInchY);
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (actualComp != 1 && actualComp != 3 && actualComp != 4)
    {
        image.error(kStrImageWrongComponents); free(decoded); return;
    }
    if (!imageIsValidSize(1, width, height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    int decodedComp = (requestedComp == -1) ? actualComp : requestedComp;
    switch (decodedComp)
    {
        case 1: image._type = PixelType.l8; break;
        case 3: image._type = PixelType.rgb8; break;
        case 4: image._type = PixelType.rgba8; break;
        default:
    }
    image._width = width;
    image._height = height;
    image._allocArea = decoded;
    image._data = decoded;
    image._pitch = width * decodedComp;
    image._pixelAspectRatio = pixelAspectRatio == -1 ? GAMUT_UNKNOWN_ASPECT_RATIO : pixelAspectRatio;
    image._resolutionY = dotsPerInchY == -1 ? GAMUT_UNKNOWN_RESOLUTION : dotsPerInchY;
    image._layoutConstraints = LAYOUT_DEFAULT;
    image._layerCount = 1;
    image._layerOffset = 0;
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces loadQOI (plugins/qoi.d:48-140).
void loadQOI_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    int requestedComp = computeRequestedImageComponents(flags);
    if (requestedComp == 0) { image.error(kStrInvalidFlags); return; }
    if (requestedComp == -1 || requestedComp == 1 || requestedComp == 2) requestedComp = 0;

    gb200_qoi_desc desc;
    ubyte* decoded = gb200_qoi_decode(buf, len, &desc, requestedComp);
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (!imageIsValidSize(1, desc.width, desc.height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    image._allocArea = decoded;
    image._data = decoded;
    image._width = desc.width;
    image._height = desc.height;
    int decodedComp = (requestedComp == 0) ? desc.channels : requestedComp;
    image._type = decodedComp == 3 ? PixelType.rgb8 : PixelType.rgba8;
    image._pitch = desc.channels * desc.width;       // sic, as plugins/qoi.d:131
    image._pixelAspectRatio = GAMUT_UNKNOWN_ASPECT_RATIO;
    image._resolutionY = GAMUT_UNKNOWN_RESOLUTION;
    image._layoutConstraints = 0;
    image._layerCount = 1;
    image._layerOffset = 0;
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces loadQOIX (plugins/qoix.d:64-146).
void loadQOIX_b200(ref Image image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    int len;
    ubyte* buf = slurp(io, handle, len, image);
    if (buf is null) return;
    scope(exit) free(buf);

    int requestedComp = computeRequestedImageComponents(flags);
    if (requestedComp == 0) { image.error(kStrInvalidFlags); return; }

    gb200_qoix_desc desc;
    int decodedToType = -1;
    ubyte* decoded = gb200_qoix_decode(buf, len, &desc, flags, &decodedToType);
    if (decoded is null) { image.error(kStrImageDecodingFailed); return; }
    if (!imageIsValidSize(1, desc.width, desc.height)) { image.error(kStrImageTooLarge); free(decoded); return; }

    image._allocArea = decoded;
    image._data = decoded;
    image._width = desc.width;
    image._height = desc.height;
    image._layoutConstraints = 0;
    image._type = cast(PixelType) decodedToType;
    image._pitch = desc.pitchBytes;
    image._pixelAspectRatio = desc.pixelAspectRatio;
    image._resolutionY = desc.resolutionY;
    image._layerCount = 1;
    image._layerOffset = 0;
    image.convertTo(applyLoadFlags(image._type, flags), cast(LayoutConstraints) flags);
}

/// Replaces saveQOIX (plugins/qoix.d:156-241) for the images the reference routes to qoiplane10_encode (10-bit
/// greyscale), qoiplane_encode (8-bit greyscale) and qoix_encode / QOI2AVG (8-bit RGB / RGBA), premultiplied variants
/// included: the stream is the reference sub-encoder's byte for byte, without the LZ4 stage (compression = 0, which is
/// what qoix_lz4_encode itself returns whenever LZ4 does not make the file smaller). 16-bit RGB(A) goes to qoi10b_encode
/// (QOI-10b) the same way. fp32 types and vertically flipped images fall through to the reference's own saveQOIX.
bool saveQOIX_b200(ref const(Image) image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    if (page != 0) return false;
    const bool plane10 = image._type == PixelType.l16 || image._type == PixelType.la16 || image._type == PixelType.lap16;
    const bool plane8 = image._type == PixelType.l8 || image._type == PixelType.la8 || image._type == PixelType.lap8;
    const bool rgb8 = image._type == PixelType.rgb8 || image._type == PixelType.rgba8 || image._type == PixelType.rgbap8;
    const bool rgb10 = image._type == PixelType.rgb16 || image._type == PixelType.rgba16 || image._type == PixelType.rgbap16;
    if (!(plane10 || plane8 || rgb8 || rgb10) || image._pitch < 0)
        return saveQOIX(image, io, handle, page, flags, data);

    gb200_qoix_desc desc;
    desc.width = image._width;
    desc.height = image._height;
    desc.pitchBytes = image._pitch;
    desc.channels = (rgb8 || rgb10) ? ((image._type == PixelType.rgb8 || image._type == PixelType.rgb16) ? 3 : 4)
                                    : (image._type == PixelType.l16 || image._type == PixelType.l8) ? 1 : 2;
    desc.bitdepth = (plane10 || rgb10) ? 10 : 8;
    desc.colorspace = (image._type == PixelType.lap16 || image._type == PixelType.lap8 || image._type == PixelType.rgbap8 || image._type == PixelType.rgbap16) ? 2 /* QOIX_SRGB_PREMUL */ : 0 /* QOIX_SRGB */;
    desc.compression = 0;
    desc.pixelAspectRatio = image._pixelAspectRatio;
    desc.resolutionY = image._resolutionY;

    int qoilen;
    ubyte* encoded = gb200_qoix_encode(image._data, &desc, &qoilen);
    if (encoded is null) return false;
    scope(exit) free(encoded);
    return qoilen == io.write(encoded, 1, qoilen, handle);
}

/// Replaces saveBMP (plugins/bmp.d:166-194): the same file as write_bmp's (codecs/bmpenc.d:25-113); the padding bytes of
/// 24-bit rows, which the reference leaves uninitialised, are zero.
bool saveBMP_b200(ref const(Image) image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    if (page != 0) return false;
    gb200_bmp_desc desc;
    desc.width = image._width;
    desc.height = image._height;
    desc.pitchBytes = image._pitch;
    desc.type = cast(int) image._type;
    desc.ppmX = image.pixelsPerMeterX();
    desc.ppmY = image.pixelsPerMeterY();
    int len;
    ubyte* encoded = gb200_bmp_encode(image._data, &desc, &len);
    if (encoded is null) return false;
    scope(exit) free(encoded);
    return len == io.write(encoded, 1, len, handle);
}

/// Replaces saveTGA (plugins/tga.d:123-149): the TGAEncoder's file (24- or 32-bit, run-length coded, bottom row first),
/// every scanline coded in parallel on the GPU. The pixel types TGAEncoder.initialize refuses are refused here too.
bool saveTGA_b200(ref const(Image) image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    if (page != 0) return false;
    gb200_tga_desc desc;
    desc.width = image._width;
    desc.height = image._height;
    desc.pitchBytes = image._pitch;
    desc.type = cast(int) image._type;
    int len;
    ubyte* encoded = gb200_tga_encode(image._data, &desc, &len);
    if (encoded is null) return false;
    scope(exit) free(encoded);
    return len == io.write(encoded, 1, len, handle);
}

/// Replaces saveQOI (plugins/qoi.d:150-185): same checks, same stream (qoi_encode's byte for byte), the encoder runs
/// on the GPU. A vertically flipped image is taken as it is (negative pitch), like the reference's qoi_encode.
bool saveQOI_b200(ref const(Image) image, IOStream* io, IOHandle handle, int page, int flags, void* data) @trusted
{
    if (page != 0) return false;
    gb200_qoi_desc desc;
    desc.width = image._width;
    desc.height = image._height;
    desc.colorspace = 0; // QOI_SRGB, as the reference (plugins/qoi.d:159)
    switch (image._type)
    {
        case PixelType.rgb8:  desc.channels = 3; break;
        case PixelType.rgba8: desc.channels = 4; break;
        default: return false; // not supported
    }
    int qoilen;
    ubyte* encoded = gb200_qoi_encode(image._data, &desc, image._pitch, &qoilen);
    if (encoded is null) return false;
    scope(exit) free(encoded);
    return qoilen == io.write(encoded, 1, qoilen, handle);
}

/// Replaces scanlinesConvert (scanline.d:70-121) for Image.convertTo (image.d:1296): same signature; the
/// reference's interType / interBuf are accepted and unused (both stages are fused on the GPU).
bool scanlinesConvert_b200(PixelType srcType, const(ubyte)* src, int srcPitch,
                           PixelType destType, ubyte* dest, int destPitch,
                           int width, int height, PixelType interType, ubyte* interBuf) @system
{
    return gb200_scanlines_convert(cast(int) srcType, src, srcPitch, cast(int) destType, dest, destPitch,
                                   width, height) != 0;
}

/// Replaces scanlinesCopy (scanline.d:37-55).
bool scanlinesCopy_b200(PixelType type, const(ubyte)* src, int srcPitch, ubyte* dest, int destPitch,
                        int width, int height) @system
{
    return gb200_scanlines_convert(cast(int) type, src, srcPitch, cast(int) type, dest, destPitch, width, height) != 0;
}
