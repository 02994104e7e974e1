// rt_scene_dev.h — device-side scene view: what the kernels read (DESIGN.md "data layout in HBM").
#pragma once
#include "../../include/rt_b200.h"
#include "rt_math.h"

// 8-wide BVH node, 128 bytes = 8 x float4 = exactly one cache line (children of a node are contiguous, Ylitie-Karras-Laine
// style, but the child planes are bfloat16 instead of 8-bit grid coordinates):
//   n0: origin.xyz (f32, = the node box's low corner) | imask << 24 | E      (imask bit i: slot i is an inner node;
//                                                                             2^(E-127) >= every plane value of the node)
//   n1: child_base (u32) | prim_base (u32) | meta[0..3] | meta[4..7]
//   n2, n3: x planes of children 0..3, 4..7   n4, n5: y planes   n6, n7: z planes
//           one 32-bit word per child and axis: (hi_bf16 << 16) | lo_bf16, planes relative to the origin (>= 0),
//           lo rounded down, hi rounded up.  Read as a float the word IS the hi plane (the lo bits underneath only enlarge
//           it by less than one bf16 ulp: conservative); word << 16 is the lo plane exactly.  So the slab test needs no
//           byte permutes / int->float conversions at all (they saturated the alu pipe with the 8-bit layout, profiles/r02):
//           an integer multiply by 1 or 65536 (fma pipe) picks the near / far plane by ray direction.
//           Empty slot: lo = +inf, hi = 0.
// meta[i]: 0 = empty; inner: 0b001_xxxxx with xxxxx = 24 + slot; leaf: top 3 bits = unary count (1 -> 001,
// 2 -> 011, 3 -> 111), low 5 bits = first primitive offset (0..23) relative to prim_base.
#define RT_NODE_F4 8
#define RT_NODE_BYTES (RT_NODE_F4 * 16)
#define RT_O2W_F4 4
#ifndef RT_LEAF_MAX
#define RT_LEAF_MAX 3
#endif
#define RT_STACK_SIZE 48

// BLAS primitive record, 48 bytes = 3 x float4: v0.xyz | primitive_id ; v1.xyz | instance_id (merged BLAS only) ; v2.xyz | 0
#define RT_TRI_F4 3

// per-instance record for traversal, 64 bytes = 4 x float4: world->object 3x4 row-major (rows 0..2),
// then { blas_root (node index), geo_id, flags (bit0: opaque geometry), 0 }
#define RT_INST_F4 4
#define RT_INST_OPAQUE 1u
// The merged world-space BLAS: instances whose geometry is referenced once (or is tiny) are baked to world space
// and share one BLAS, entered through a pseudo instance record with an identity transform (no ray transform, the
// instance id comes from the triangle record).  DESIGN.md "baked instances".
#define RT_INST_IDENTITY 2u
#define RT_INST_MERGED 4u
#define RT_BAKE_MAX_TRIS 256u

struct DImage { const uint8_t* px; uint32_t w, h, srgb, _pad; };
struct DTexture { uint32_t image, mag_filter, wrap_s, wrap_t; };

struct DScene {
    // traversal
    const float4* tlas_nodes;      // RT_NODE_F4 float4 per node, root = 0
    const uint32_t* tlas_prims;    // instance ids in TLAS leaf order
    const float4* blas_nodes;      // all BLASes, child indices absolute
    const float4* tris;            // RT_TRI_F4 float4 per triangle, leaf order
    const float4* inst_w2o;        // RT_INST_F4 float4 per instance
    const float4* inst_o2w;        // RT_O2W_F4 float4 per instance: object->world rows | v_offset, i_offset, material_id, geo_id
    uint32_t n_instances;
    // when the TLAS holds nothing but the merged world-space BLAS, rays start inside it (no TLAS visit, no instance entry)
    uint32_t single_merged, merged_node_off, merged_tri_off;
    // merged BLAS next to real instances: rays also start inside it (it is not a TLAS entry) and continue at the TLAS
    // root when it is exhausted — one instance entry less per ray, the world-space shear is set up with the ray
    uint32_t merged_first;
    // shading inputs in the reference's layouts
    const rt_vertex* vertices;     // skinned output (AnimationCompute.comp) == BLAS build input
    const uint32_t* indices;
    const rt_prim_info* prim_infos;
    const rt_material* materials;
    const DTexture* textures; uint32_t n_textures;
    const DImage* images;
    const rt_light* dlights; uint32_t n_dlights;
    const rt_light* plights; uint32_t n_plights;
    uint32_t nee_plights;          // n_plights if a point light of the first min(n,3) slots is bright enough for RIS (RayTracing.rchit:86), else 0
    const float* srgb_lut;         // 256 entries
    DImage sky[6]; uint32_t has_sky_faces;
};

struct RtCounters { unsigned long long nodes, tris, insts, anyhits, tex_taps, light_cands; };

// Wavefront queues (SoA): one path = 64 B of state
struct DQueue {
    float4* o_tmin;     // origin.xyz, tmin
    float4* d_tmax;     // direction.xyz, tmax
    float4* thr_pix;    // throughput.xyz, pixel index (uint bits)
    uint4*  rng;        // PATH.w, PIX.w, LENS seed (LCG), volume_dis (float bits)
};
struct DHits { float4* tuvp; uint32_t* inst; };   // t,u,v,primitive(bits) | instance
struct DShadowQueue {
    float4* o_tmax;     // origin.xyz, tmax
    float4* d_pix;      // direction.xyz, pixel index (uint bits)
    float4* contrib;    // throughput * brdf * weight * intensity * colour, PATH.w at chit entry (uint bits)
};
