// astc_save.h -- .astc container writer and GPU read-back, header-only like the
// reference's astc_save.h, over the C ABI of astc_b200.h.
//
//   astc_header            reference astc_save.h:5-14   (same 16-byte layout)
//   read_gpu(...)          reference astc_save.h:34-50  (staging copy + Map -> D2H copy + sync)
//   save_astc(...)         reference astc_save.h:52-76  (same argument list)
#pragma once
#include <cstdint>
#include <cstdio>

#include "astc_cuda_handles.h"

#define MAGIC_FILE_CONSTANT 0x5CA1AB13

// 16 bytes on disk: magic (LE), block dims, 24-bit LE texel sizes.
struct astc_header {
    uint8_t magic[4];
    uint8_t blockdim_x;
    uint8_t blockdim_y;
    uint8_t blockdim_z;
    uint8_t xsize[3];
    uint8_t ysize[3];
    uint8_t zsize[3];
};
static_assert(sizeof(astc_header) == 16, "astc_header is a packed 16-byte record");

namespace astc_save {

inline void put24(uint8_t dst[3], int v)
{
    for (int i = 0; i < 3; ++i) dst[i] = uint8_t((v >> (8 * i)) & 0xFF);
}

inline astc_header make_header(int xdim, int ydim, int xsize, int ysize)
{
    astc_header h{};
    for (int i = 0; i < 4; ++i) h.magic[i] = uint8_t((uint32_t(MAGIC_FILE_CONSTANT) >> (8 * i)) & 0xFF);
    h.blockdim_x = uint8_t(xdim);
    h.blockdim_y = uint8_t(ydim);
    h.blockdim_z = 1;
    put24(h.xsize, xsize);
    put24(h.ysize, ysize);
    put24(h.zsize, 1);
    return h;
}

// Unlike the reference (which would crash on a null FILE*), reports failure.
inline bool write_file(const char *path, int xdim, int ydim, int xsize, int ysize, const uint8_t *blocks, size_t bufsz)
{
    std::FILE *f = std::fopen(path, "wb");
    if (!f) return false;
    const astc_header h = make_header(xdim, ydim, xsize, ysize);
    bool ok = std::fwrite(&h, 1, sizeof h, f) == sizeof h;
    ok = ok && (bufsz == 0 || std::fwrite(blocks, 1, bufsz, f) == bufsz);
    ok = (std::fclose(f) == 0) && ok;
    return ok;
}

}  // namespace astc_save

// Download the encoder's output buffer; returns 0 (S_OK) or a negative status.
inline int read_gpu(astc_device * /*pDevice*/, astc_context *pContext, astc_buffer *pBuffer, uint8_t *pMemBuf, uint32_t buf_len)
{
    if (!pBuffer || !pMemBuf) return ASTC_B200_ERR_INVALID_ARGUMENT;
    void *stream = pContext ? pContext->stream : nullptr;
    int rc = astc_b200_memcpy_d2h(pMemBuf, pBuffer->d_data, buf_len, stream);
    if (rc == ASTC_B200_OK) rc = astc_b200_stream_synchronize(stream);
    return rc;
}

inline void save_astc(const char *astc_path, int xdim, int ydim, int xsize, int ysize, uint8_t *buffer, int bufsz)
{
    astc_save::write_file(astc_path, xdim, ydim, xsize, ysize, buffer, size_t(bufsz));
}
