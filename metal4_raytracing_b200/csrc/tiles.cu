// tiles.cu — screen-tile slabs for the multi-GPU frame exchange.
//
// The reference is single-device (MetalRaytracing/Renderer.swift:229); its 16x16 threadgroup tiling
// (Renderer.swift:1445-1451) is the unit of ownership here: rank g renders tiles with tile % N == g. When the
// frame is assembled with an NCCL all-gather instead of direct peer stores (trace.cu peerAccumulation), each rank
// first packs its owned tiles into a dense slab [ownedTile][16*16][bytesPerPixel]; after the gather every rank
// scatters the N slabs back into a full frame. Both kernels move whole pixels with the widest aligned access the
// pixel size allows and are pure HBM streaming.
#include "common.cuh"

namespace rtb {

template <typename Pixel>
__global__ void k_pack_tiles(const Pixel *__restrict__ image, Pixel *__restrict__ slab, int width, int height,
                             int tilesX, int tileCount, int modulo, int remainder) {
  const int owned = blockIdx.x;
  const int tile = owned * modulo + remainder;
  const int x = (tile % tilesX) * 16 + (threadIdx.x & 15), y = (tile / tilesX) * 16 + (threadIdx.x >> 4);
  Pixel v{}; // slab padding (a rank with one tile fewer, pixels beyond a ragged edge) is zero
  if (tile < tileCount && x < width && y < height) v = image[size_t(y) * width + x];
  slab[size_t(owned) * 256 + threadIdx.x] = v;
}

template <typename Pixel>
__global__ void k_unpack_tiles(const Pixel *__restrict__ slabs, Pixel *__restrict__ image, int width, int height,
                               int tilesX, int tileCount, int modulo, int slabTiles) {
  const int tile = blockIdx.x;
  if (tile >= tileCount) return;
  const int rank = tile % modulo, owned = tile / modulo;
  const int x = (tile % tilesX) * 16 + (threadIdx.x & 15), y = (tile / tilesX) * 16 + (threadIdx.x >> 4);
  if (x < width && y < height)
    image[size_t(y) * width + x] = slabs[(size_t(rank) * slabTiles + owned) * 256 + threadIdx.x];
}

static int pixelBytes(int format) {
  switch (format) {
    case RT_FORMAT_R32_UINT:
    case RT_FORMAT_R32_FLOAT:
    case RT_FORMAT_RG16_FLOAT: return 4;
    case RT_FORMAT_RGBA16_FLOAT:
    case RT_FORMAT_RG32_FLOAT: return 8;
    case RT_FORMAT_R16_FLOAT: return 2;
    case RT_FORMAT_RGBA32_FLOAT: return 16;
    default: return 0;
  }
}

int packTiles(rt_context *ctx, const rt_image *image, void *slab, int modulo, int remainder) {
  RT_CHECK(image && image->data && slab, "rt_pack_tiles: null pointer");
  RT_CHECK(modulo >= 1 && remainder >= 0 && remainder < modulo, "rt_pack_tiles: bad tile partition");
  const int tilesX = (image->width + 15) / 16, tilesY = (image->height + 15) / 16, tileCount = tilesX * tilesY;
  const int slabTiles = (tileCount + modulo - 1) / modulo;
  const int bytes = pixelBytes(image->format);
  RT_CHECK(bytes != 0, "rt_pack_tiles: unknown image format");
#define RT_PACK(T)                                                                                                   \
  k_pack_tiles<T><<<slabTiles, 256, 0, ctx->stream>>>(static_cast<const T *>(image->data), static_cast<T *>(slab),   \
                                                      image->width, image->height, tilesX, tileCount, modulo, remainder)
  if (bytes == 16) RT_PACK(uint4);
  else if (bytes == 8) RT_PACK(uint2);
  else if (bytes == 4) RT_PACK(uint32_t);
  else RT_PACK(uint16_t);
#undef RT_PACK
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

int unpackTiles(rt_context *ctx, const void *slabs, const rt_image *image, int modulo) {
  RT_CHECK(image && image->data && slabs, "rt_unpack_tiles: null pointer");
  RT_CHECK(modulo >= 1, "rt_unpack_tiles: bad tile partition");
  const int tilesX = (image->width + 15) / 16, tilesY = (image->height + 15) / 16, tileCount = tilesX * tilesY;
  const int slabTiles = (tileCount + modulo - 1) / modulo;
  const int bytes = pixelBytes(image->format);
  RT_CHECK(bytes != 0, "rt_unpack_tiles: unknown image format");
#define RT_UNPACK(T)                                                                                                 \
  k_unpack_tiles<T><<<tileCount, 256, 0, ctx->stream>>>(static_cast<const T *>(slabs), static_cast<T *>(image->data), \
                                                        image->width, image->height, tilesX, tileCount, modulo, slabTiles)
  if (bytes == 16) RT_UNPACK(uint4);
  else if (bytes == 8) RT_UNPACK(uint2);
  else if (bytes == 4) RT_UNPACK(uint32_t);
  else RT_UNPACK(uint16_t);
#undef RT_UNPACK
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
