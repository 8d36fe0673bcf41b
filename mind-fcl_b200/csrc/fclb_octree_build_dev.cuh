// fclb_octree_build_dev.cuh -- octree2::Octree<S>::rebuildTree on the DEVICE, with the reference's node numbering.
//
// Reference: geometry/octree2/octree-inl.h:15-142 (layers, root box, computeVoxelCoordinate, computeChildIndex),
// octree_construction-inl.h:10-74 (insertVoxelIntoTree), :111-205 (fully-occupied flags, rebuildTree).
// The reference inserts the points one by one and APPENDS a node the first time the stream reaches it; the octree
// kernels report contacts by those node indices (encodeOctree2Node), so the numbering is part of the contract.  It is a
// pure function of the stream: a node's creation time is (index of the first point that reaches it, its depth) --
// a point creates the missing nodes of its path top-down -- so
//   inner node index = 1 + rank of (first point, depth) among the inner nodes below the root,
//   leaf  node index = rank of (first point) among the leaf-layer nodes.
// The device therefore sorts the voxel path keys (cub radix sort, stable: the first entry of a key is its first
// point), folds them level by level into unique prefixes with the minimum first-point index, ranks the nodes by their
// creation time with a second sort, and links children / ORs the leaf masks / derives the fully-occupied flags with
// one small kernel per level.  tests/test_octree_build_gpu.py: every array equals the reference's.
#pragma once
#include <cstdint>

namespace fclb {

// path key of a point: child index at every layer from the root (3 bits each, parent layer 0 first), or ~0 when the
// point falls outside the grid.  Voxel coordinate as Octree<S>::computeVoxelCoordinate: int(floor(p * inv) + S(half)).
template <typename S>
__global__ void octKeyKernel(const S* __restrict__ pts, size_t n, S inv, int half, int num_layers, unsigned long long* __restrict__ key,
                             uint32_t* __restrict__ idx) {
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const S px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    unsigned long long k = ~0ull;
    if (px == px && py == py && pz == pz) {
      const S fx = floor(px * inv) + S(half), fy = floor(py * inv) + S(half), fz = floor(pz * inv) + S(half);
      const S lim = S(2 * half);
      // (the reference truncates to int first; a value outside [0, 2 half) fails its range check either way)
      if (fx > S(-1) && fx < lim && fy > S(-1) && fy < lim && fz > S(-1) && fz < lim) {
        const int x = int(fx), y = int(fy), z = int(fz);
        if (x >= 0 && y >= 0 && z >= 0) {
          k = 0;
          #pragma unroll 1
          for (int layer = 0; layer <= num_layers - 2; layer++) {  // computeChildIndex(voxel, parent layer)
            const int diff = num_layers - layer - 2;
            const unsigned c = ((x >> diff) & 1) | (((y >> diff) & 1) << 1) | (((z >> diff) & 1) << 2);
            k = (k << 3) | c;
          }
        }
      }
    }
    key[i] = k;
    idx[i] = uint32_t(i);
  }
}

// items of one level: sorted unique prefixes (key), first point (first), position of the parent at the level above
struct OctLevel {
  unsigned long long* key = nullptr;
  uint32_t* first = nullptr;
  uint32_t* parent = nullptr;  // position in the level above
  uint32_t* index = nullptr;   // reference node index of the item (filled by the ranking pass)
  uint32_t n = 0;
};

// head[j] = 1 when item j starts a new prefix (key >> shift differs from its predecessor's)
__global__ void octHeadKernel(const unsigned long long* __restrict__ key, uint32_t n, int shift, uint32_t* __restrict__ head) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    head[j] = (j == 0 || (key[j] >> shift) != (key[j - 1] >> shift)) ? 1u : 0u;
}
// scanned[j] = inclusive scan of head => parent position = scanned[j] - 1
__global__ void octFoldKernel(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ first, const uint32_t* __restrict__ scanned,
                              uint32_t n, int shift, unsigned long long* __restrict__ pkey, uint32_t* __restrict__ pfirst,
                              uint32_t* __restrict__ parent) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t p = scanned[j] - 1;
    parent[j] = p;
    pkey[p] = key[j] >> shift;  // (every member of the segment writes the same value)
    atomicMin(&pfirst[p], first[j]);
  }
}
__global__ void octFillKernel(uint32_t* p, uint32_t n, uint32_t v) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) p[j] = v;
}
// creation-time key of the inner nodes of one level: (first point << 5) | depth, value = (level, position) packed
__global__ void octRankKeyKernel(const uint32_t* __restrict__ first, uint32_t n, int depth, uint32_t offset, unsigned long long* __restrict__ key,
                                 uint32_t* __restrict__ val) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    key[offset + j] = (static_cast<unsigned long long>(first[j]) << 5) | unsigned(depth);
    val[offset + j] = offset + j;
  }
}
// sorted_val[r] = flat position of the node with rank r  =>  index[flat position] = r + base
__global__ void octAssignIndexKernel(const uint32_t* __restrict__ sorted_val, uint32_t n, uint32_t base, uint32_t* __restrict__ flat_index) {
  #pragma unroll 1
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) flat_index[sorted_val[r]] = r + base;
}
// children[8 * index(parent) + c] = index(item), c = the item's last 3 key bits
__global__ void octLinkKernel(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ parent, const uint32_t* __restrict__ index,
                              uint32_t n, const uint32_t* __restrict__ parent_index, uint32_t* __restrict__ children) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t pi = parent_index ? parent_index[parent[j]] : 0u;
    children[size_t(8) * pi + unsigned(key[j] & 7ull)] = index[j];
  }
}
// voxels (full keys) -> the occupancy mask of their leaf node
__global__ void octLeafBitsKernel(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ parent, uint32_t n,
                                  const uint32_t* __restrict__ leaf_index, uint32_t* __restrict__ bits32) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    atomicOr(&bits32[leaf_index[parent[j]]], 1u << unsigned(key[j] & 7ull));
}
__global__ void octNarrowKernel(const uint32_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) out[j] = uint8_t(in[j]);
}
// fully-occupied flag of the inner nodes of one depth (octree_construction-inl.h:111-172): all eight children present
// and full (a leaf-layer child: mask 0xff)
__global__ void octFullKernel(const uint32_t* __restrict__ index, uint32_t n, const uint32_t* __restrict__ children, int children_are_leaves,
                              const uint8_t* __restrict__ leaf_bits, uint8_t* __restrict__ full) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t node = index ? index[j] : 0u;
    bool all = true;
    for (int c = 0; c < 8; c++) {
      const uint32_t ch = children[size_t(8) * node + c];
      if (ch == 0xffffffffu || (children_are_leaves ? leaf_bits[ch] != 0xff : full[ch] == 0)) all = false;
    }
    full[node] = all ? 1 : 0;
  }
}
__global__ void octUniqueKernel(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ scanned,
                                uint32_t n, unsigned long long* __restrict__ ukey, uint32_t* __restrict__ ufirst) {
  #pragma unroll 1
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    if (j == 0 || key[j] != key[j - 1]) {  // stable sort: the first entry of a key carries its first point
      ukey[scanned[j] - 1] = key[j];
      ufirst[scanned[j] - 1] = idx[j];
    }
}

}  // namespace fclb
