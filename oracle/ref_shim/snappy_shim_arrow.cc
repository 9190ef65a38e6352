// oracle/ref_shim/snappy_shim_arrow.cc -- TEST INFRASTRUCTURE ONLY.
// snappy-c.h provider over the genuine Google Snappy that pyarrow bundles in libarrow
// (arrow::util::Codec, Compression::SNAPPY = raw Snappy block format).  SURVEY.md Appendix E.
#include "snappy-c.h"
#include <arrow/util/compression.h>
#include <cstdint>
#include <memory>

static arrow::util::Codec *codec()
{
    static std::unique_ptr<arrow::util::Codec> c = *arrow::util::Codec::Create(arrow::Compression::SNAPPY);
    return c.get();
}

extern "C" size_t snappy_max_compressed_length(size_t n) { return 32 + n + n / 6; }

extern "C" snappy_status snappy_uncompressed_length(const char *in, size_t n, size_t *result)
{
    uint64_t v = 0;
    for (size_t i = 0; i < n && i < 5; i++) {
        uint8_t b = (uint8_t)in[i];
        v |= (uint64_t)(b & 0x7F) << (7 * i);
        if (!(b & 0x80)) {
            if (v > 0xFFFFFFFFull) return SNAPPY_INVALID_INPUT;
            *result = (size_t)v;
            return SNAPPY_OK;
        }
    }
    return SNAPPY_INVALID_INPUT;
}

extern "C" snappy_status snappy_compress(const char *in, size_t n, char *out, size_t *out_len)
{
    if (*out_len < snappy_max_compressed_length(n)) return SNAPPY_BUFFER_TOO_SMALL;
    auto r = codec()->Compress((int64_t)n, (const uint8_t *)in, (int64_t)*out_len, (uint8_t *)out);
    if (!r.ok()) return SNAPPY_INVALID_INPUT;
    *out_len = (size_t)*r;
    return SNAPPY_OK;
}

extern "C" snappy_status snappy_uncompress(const char *in, size_t n, char *out, size_t *out_len)
{
    size_t want;
    if (snappy_uncompressed_length(in, n, &want) != SNAPPY_OK) return SNAPPY_INVALID_INPUT;
    if (*out_len < want) return SNAPPY_BUFFER_TOO_SMALL;
    auto r = codec()->Decompress((int64_t)n, (const uint8_t *)in, (int64_t)want, (uint8_t *)out);
    if (!r.ok() || (size_t)*r != want) return SNAPPY_INVALID_INPUT;
    *out_len = want;
    return SNAPPY_OK;
}
