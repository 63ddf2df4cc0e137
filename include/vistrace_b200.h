/*
 * vistrace_b200.h — C ABI of the B200-native ray-query engine that sits behind
 * VisTrace's `vistrace.CreateAccel` / `accel:Traverse` surface.
 *
 * Plain C: pointers, sizes and POD records only.  No STL, no torch types, no
 * exceptions cross this boundary.  Every entry point returns 0 on success (or
 * a handle / NULL) and leaves a message for vt_last_error() on failure.
 *
 * Each declaration cites the reference interface it replaces as
 * `path:line` relative to the Derpius/VisTrace tree.
 */
#ifndef VISTRACE_B200_H
#define VISTRACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_ABI_VERSION 1u
#define VT_MISS 0xFFFFFFFFu /* vt_hit.prim for "Traverse returned nil" (source/objects/AccelStruct.cpp:837) */

/* ------------------------------------------------------------------ records */

/* One ray.  Replaces bvh::Ray<float> (source/objects/Primitives.h:9-33) minus the
 * pAccel back-pointer.  The direction is NOT normalised for traversal, t is in
 * units of |d| (source/objects/AccelStruct.cpp:810-815). 32 bytes. */
typedef struct vt_ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} vt_ray;

/* Closest hit.  Replaces ClosestPrimitiveIntersector::Result
 * (libs/bvh/include/bvh/primitive_intersectors.hpp:48-53) +
 * TriangleBackfaceCull::Intersection (source/objects/Primitives.h:46-50).
 * prim indexes the ORIGINAL triangle array; VT_MISS when nothing was hit
 * (then t = u = v = 0). 16 bytes. */
typedef struct vt_hit {
    float t, u, v;
    uint32_t prim;
} vt_hit;

/* BVH node, bit-compatible with bvh::Bvh<float>::Node
 * (libs/bvh/include/bvh/bvh.hpp:25-31): bounds = {minx,maxx,miny,maxy,minz,maxz},
 * prim_count != 0 marks a leaf, `first` is the first child (children are
 * adjacent: first, first+1) or the first slot of prim_indices. 32 bytes. */
typedef struct vt_node {
    float bounds[6];
    uint32_t prim_count;
    uint32_t first;
} vt_node;

/* Input triangle: what mesh ingestion hands to the Triangle constructor
 * (source/objects/Primitives.h:75-89, source/objects/Model.cpp:82-89) plus the
 * per-vertex attributes the ingestion code fills in afterwards. 152 bytes. */
typedef struct vt_tri_in {
    float p[3][3];        /* p0, p1, p2 */
    float normals[3][3];  /* Triangle::normals */
    float tangents[3][3]; /* Triangle::tangents */
    float uvs[3][2];      /* Triangle::uvs */
    float alphas[3];      /* Triangle::alphas */
    uint32_t material;    /* global material index (Triangle::material) */
    uint16_t ent_idx;     /* index into the entity array (Triangle::entIdx) */
    uint8_t one_sided;    /* Triangle::oneSided */
    uint8_t pad;
} vt_tri_in;

/* Per-triangle skinning data, the fields Model.cpp leaves in a Triangle for SkinTriangle
 * (source/objects/Primitives.h:68-70): per vertex up to three (bone, weight) influences. 52 bytes. */
typedef struct vt_tri_skin {
    uint8_t num_bones[3];   /* Triangle::numBones */
    uint8_t pad;
    int8_t bone_ids[3][3];  /* Triangle::boneIds */
    uint8_t pad2[3];
    float weights[3][3];    /* Triangle::weights */
} vt_tri_skin;

/* Texture: decoded mip chain in the order VTFTexture keeps it in memory, i.e. SMALLEST mip first
 * (libs/VTFParser/VTFParser.cpp:44-78,178-188), one frame, one face, depth 1.  flags are VTF TEXTURE_FLAGS (CLAMPS 0x4, CLAMPT 0x8).
 * texel_layout = 0: RGBA8888, a channel is byte / 255.f (ParsePixel's RGBA8888 case, FileFormat/Parser.cpp:161-168).
 * texel_layout = VT_TEXEL_WIDE | codes: 8 bytes per texel, four uint16 NUMERATORS r, g, b, a; channel c is numerator / divisor with
 * the divisor named by the 2-bit code in bits 2c, 2c + 1 (VT_TEXEL_DIV_255 / _65535 / _1) — what the reference's 16-bit formats need:
 * their ParsePixel expressions exceed 255 / 255 (Parser.cpp:190-196,238-262) or divide by 65535 (:281-294).  vt_vtf_decode emits it. */
#define VT_TEXFLAG_CLAMPS 0x00000004u
#define VT_TEXFLAG_CLAMPT 0x00000008u
#define VT_TEXEL_WIDE      0x100u
#define VT_TEXEL_DIV_255   0u
#define VT_TEXEL_DIV_65535 1u
#define VT_TEXEL_DIV_1     2u
typedef struct vt_texture {
    uint16_t width, height; /* size of mip 0 */
    uint16_t mip_count;
    uint16_t pad;
    uint32_t flags;
    uint32_t texel_layout;
    const uint8_t *rgba;    /* host pointer, nbytes long: 4 (or, wide, 8) bytes per texel */
    uint64_t nbytes;
} vt_texture;

/* VTF file -> the mip chain of a vt_texture; host only.  Restates libs/VTFParser: header and image-data
 * location (FileFormat/Parser.cpp:99-155, FileFormat/Structs.h:21-75), DXT1/3/5 decompressed at load
 * (VTFParser.cpp:26-84, DXTn/DXT1.cpp, DXT3.cpp, DXT5.cpp), every other format converted as ParsePixel reads it
 * (Parser.cpp:157-298): the 8-bit-per-channel formats (and P8, which ParsePixel reads as opaque black) to RGBA8888, the 16-bit
 * formats (RGB565, BGR565, BGRX5551, BGRA5551, BGRA4444, RGBA16161616(F)) to WIDE texels (vt_texture.texel_layout) — every channel
 * the device samples equals the float VTFTexture::GetPixel returns, bit for bit. */
typedef struct vt_vtf_info {
    uint32_t width, height; /* mip 0 */
    uint32_t mip_count;
    uint32_t flags;         /* VTF TEXTURE_FLAGS: pass on as vt_texture.flags (CLAMPS / CLAMPT) */
    int32_t format;         /* IMAGE_FORMAT of the file (FileFormat/Enums.h:5-35) */
    uint32_t frames, faces, depth;
    uint32_t supported;     /* 1 when vt_vtf_decode can produce the chain (every format ParsePixel knows) */
    uint32_t texel_layout;  /* pass on as vt_texture.texel_layout: 0 = RGBA8888, else wide texels */
    uint64_t rgba_bytes;    /* size of the decoded chain of ONE frame / face / z-slice 0, smallest mip first */
} vt_vtf_info;
int vt_vtf_read_info(const uint8_t *file, uint64_t size, vt_vtf_info *info);
/* rgba_out: capacity >= info.rgba_bytes; the chain is in VTF order (smallest mip first), ready for vt_texture.rgba. */
int vt_vtf_decode(const uint8_t *file, uint64_t size, uint32_t frame, uint32_t face, uint8_t *rgba_out, uint64_t capacity,
                  vt_vtf_info *info_or_null);

/* Source-engine model ingestion, host only: .mdl + .vvd + .vtx -> the triangles the reference's Model /
 * Mesh classes hand to PopulateAccel.  Restates libs/MDLParser (MDLParser.cpp:33-60, VVDParser.cpp:33-83, VTXParser.cpp:33-56:
 * ids, versions <= 48 / 4 / 7, the checksum tying the three files together, the VVD fix-up table) and source/objects/Model.cpp
 * (Mesh::Mesh :11-128: LoD 0, triangle-list strips, normals / tangents through glm::normalize, a non-finite tangent replaced by
 * normalize(e1), per-vertex bone weights; bind matrices :242-254; skin table :349-357; material paths :256-274).  Every offset and
 * index is checked against the file sizes.  The triangles come out in MODEL space with their skinning data: vt_skin_triangles
 * (SkinTriangle) bakes them into world space, as AccelStruct::PopulateAccel does per entity (source/objects/AccelStruct.cpp:716-749). */
typedef struct vt_mdl_files {
    const uint8_t *mdl; uint64_t mdl_size;
    const uint8_t *vvd; uint64_t vvd_size;
    const uint8_t *vtx; uint64_t vtx_size;
} vt_mdl_files;
typedef struct vt_mdl_info {
    uint32_t version;         /* studio header version (<= 48) */
    uint32_t n_bodygroups;    /* Model::GetNumBodyGroups */
    uint32_t n_bones;         /* Model::GetNumBones */
    uint32_t n_materials;     /* Model::GetNumMaterials */
    uint32_t n_material_dirs;
    uint32_t n_skin_refs, n_skin_families; /* Model::GetNumSkinFamilies */
    uint32_t n_vertices;      /* root-LoD vertices of the vvd */
} vt_mdl_info;
int vt_mdl_read_info(const vt_mdl_files *files, vt_mdl_info *info);
int vt_mdl_bodygroup_values(const vt_mdl_files *files, uint32_t bodygroup, uint32_t *n_values); /* BodyGroup::GetNumMeshes */
/* Model::GetMesh(bodygroup, value)->GetTriangles(): call with tris == NULL for the count; *n_tris = capacity in, count out.
 * tris[i].material is the MODEL-LOCAL material id (Mesh::material); map it with vt_mdl_material_index. */
int vt_mdl_mesh_triangles(const vt_mdl_files *files, uint32_t bodygroup, uint32_t value, vt_tri_in *tris, vt_tri_skin *skin, uint64_t *n_tris);
int vt_mdl_bind_matrices(const vt_mdl_files *files, float *out16); /* Model::GetBindMatrix: n_bones glm::mat4, column-major */
int vt_mdl_material_index(const vt_mdl_files *files, uint32_t skin, uint32_t material_id, int32_t *index); /* Model::GetMaterialIdx */
int vt_mdl_material_path(const vt_mdl_files *files, uint32_t material_id, uint32_t dir, char *out, uint64_t capacity); /* directory + name */

/* Source-engine map ingestion, host only: .bsp (VBSP 19-21) -> the world triangles, materials and static-prop
 * placements the reference's World object is built from.  Restates libs/BSPParser (FileFormat/Parser.cpp:11-73: which lumps a valid
 * map carries; BSPParser.cpp:184-523 Triangulate: worldspawn faces, nodraw / skip / trigger dropped, polygons fanned from their first
 * vertex, flat normals and tangent frames from the texture axes, UVs from the texture vectors; Displacements/ (all six files):
 * displacement vertices, normals, tangent frames, UVs and the three smoothing passes across neighbouring displacements) and the
 * world half of World::World (source/objects/AccelStruct.cpp:236-414: one material per distinct texture path in order of first use,
 * every world triangle one-sided, entity 0).  The records equal BSPMap's arrays bit for bit.  Every offset, count and index is
 * checked against the file: where the reference would read out of bounds the file is rejected. */
typedef struct vt_bsp_info {
    uint32_t version;              /* 19, 20 or 21 */
    uint32_t n_materials;          /* distinct texture paths among the emitted triangles, in order of first use */
    uint32_t n_texinfos;
    uint32_t n_displacements;
    uint32_t n_static_props;       /* BSPMap::GetNumStaticProps */
    uint32_t static_props_version; /* 4, 5 or 6; 0 when the map has no static-prop lump */
    uint64_t n_tris;               /* BSPMap::GetNumTris */
} vt_bsp_info;
typedef struct vt_bsp_material {
    uint32_t surf_flags;   /* BSPTexture::flags (BSPEnums::SURF) of the texinfo that introduced the path -> vt_material.surf_flags */
    int32_t texinfo;       /* that texinfo */
    int32_t width, height; /* BSPTexture::width / height */
    float reflectivity[3]; /* BSPTexture::reflectivity */
    char path[260];        /* BSPTexture::path: the name World::World hands to Material() */
} vt_bsp_material;
typedef struct vt_bsp_static_prop {
    float pos[3];   /* BSPStaticProp::pos */
    float ang[3];   /* BSPStaticProp::ang (QAngle: pitch, yaw, roll) */
    int32_t skin;   /* BSPStaticProp::skin */
    char model[128]; /* BSPStaticProp::model: the .mdl path (vt_mdl_* ingests it) */
} vt_bsp_static_prop;
int vt_bsp_read_info(const uint8_t *file, uint64_t size, vt_bsp_info *info);
/* BSPMap::GetVertices / GetNormals / GetTangents / GetUVs / GetAlphas as World::World packs them into Triangles: call with tris == NULL
 * for the count; *n_tris = capacity in, count out.  tris[i].material indexes the material list (vt_bsp_get_material), ent_idx = 0,
 * one_sided = 1.  binormals_or_null: 9 floats per triangle (GetBinormals: the reference computes them, Triangle does not keep them);
 * texinfo_or_null: one int16 per triangle (GetTriTextures). */
int vt_bsp_triangles(const uint8_t *file, uint64_t size, vt_tri_in *tris, float *binormals_or_null, int16_t *texinfo_or_null, uint64_t *n_tris);
int vt_bsp_get_material(const uint8_t *file, uint64_t size, uint32_t material, vt_bsp_material *out); /* BSPMap::GetTexture of material's first texinfo */
int vt_bsp_get_static_prop(const uint8_t *file, uint64_t size, uint32_t index, vt_bsp_static_prop *out); /* BSPMap::GetStaticProp */

/* Material subset on the path (source/objects/Material.h:74-125).  Texture slots
 * are indices into the texture array, -1 = nullptr.  *_mat are glm::mat2x4 in
 * memory order: [0..3] = column 0 (drives u), [4..7] = column 1 (drives v)
 * (source/Utils.h:65-72). */
#define VT_MATFLAG_ALPHATEST 256u  /* MaterialFlags::alphatest */
#define VT_MATFLAG_NOCULL    8192u /* MaterialFlags::nocull */
#define VT_SURF_SKY          0x4u  /* BSPEnums::SURF::SKY */
typedef struct vt_material {
    uint32_t flags;       /* MaterialFlags */
    uint32_t surf_flags;  /* BSPEnums::SURF */
    float alphatest_reference;
    float tex_scale;
    float colour[4];
    float base_tex_mat[8];
    float base_tex_mat2[8];
    float normal_map_mat[8];
    float normal_map_mat2[8];
    float blend_tex_mat[8];
    float detail_mat[8];
    float detail_scale;
    float detail_blend_factor;
    float detail_tint[3];
    int32_t base_texture;
    int32_t base_texture2;
    int32_t normal_map;
    int32_t normal_map2;
    int32_t mrao;
    int32_t mrao2;
    int32_t blend_texture;
    int32_t detail;
    uint8_t detail_blend_mode; /* DetailBlendMode */
    uint8_t masked_blending;
    uint8_t detail_alpha_mask_base_texture;
    uint8_t water;
} vt_material;

/* Entity subset on the path (source/objects/AccelStruct.h:33-40). */
typedef struct vt_entity {
    uint32_t id;     /* Entity::id, what TraceResult::entIdx reports */
    float colour[4]; /* Entity::colour */
} vt_entity;

/* Headless scene: the three containers AccelStruct owns after ingestion
 * (source/objects/AccelStruct.h:72-77).  All host pointers, borrowed for the
 * duration of the call. */
typedef struct vt_scene {
    const vt_tri_in *tris;
    uint64_t n_tris;
    const vt_material *materials;
    uint32_t n_materials;
    const vt_entity *entities;
    uint32_t n_entities;
    const vt_texture *textures;
    uint32_t n_textures;
} vt_scene;

/* Eager, batched TraceResult (source/objects/TraceResult.h:54-111): everything the
 * VisTraceResult getters return for the default (no LOD cone) call. 128 bytes. */
#define VT_ATTR_FRONT_FACING 1u
#define VT_ATTR_HIT_SKY      2u
#define VT_ATTR_HIT_WATER    4u
typedef struct vt_attr {
    float pos[3];              /* TraceResult::GetPos            TraceResult.cpp:255-262 */
    float distance;            /* TraceResult::distance (units of |dir|) */
    float normal[3];           /* GetNormal   (CalcTBN)          TraceResult.cpp:132-187 */
    float alpha;               /* GetAlpha    (CalcShadingData)  TraceResult.cpp:189-253 */
    float tangent[3];          /* GetTangent */
    float metalness;           /* GetMetalness */
    float binormal[3];         /* GetBinormal */
    float roughness;           /* GetRoughness */
    float geometric_normal[3]; /* TraceResult::geometricNormal */
    float base_mip;            /* GetBaseMIPLevel */
    float albedo[3];           /* GetAlbedo */
    uint32_t ent_id;           /* TraceResult::entIdx = Entity::id */
    float uvw[3];              /* TraceResult::uvw = (u, v, 1-u-v) */
    uint32_t submat_idx;       /* TraceResult::submatIdx = Triangle::material */
    float tex_uv[2];           /* TraceResult::texUV */
    uint32_t flags;            /* VT_ATTR_* */
    uint32_t prim;             /* original triangle index, VT_MISS on a miss (rest zero) */
} vt_attr;

/* One BSDF sample: BSDFSample (source/libraries/BSDF.h:112-118).  32 bytes. */
#define VT_LOBE_NONE               0u
#define VT_LOBE_DIFFUSE_REFLECTION 1u /* LobeType::DiffuseReflection (BSDF.h:13) */
typedef struct vt_bsdf_sample {
    float scattered[3]; /* world-space direction of the scattered ray */
    float pdf;
    float weight[3];    /* BSDF * cos / pdf of the chosen lobe, over the lobe's selection probability */
    uint32_t lobe;      /* VT_LOBE_* */
} vt_bsdf_sample;

/* --------------------------------------------------------------- entry points */

typedef struct vt_accel vt_accel; /* opaque; replaces AccelStruct (source/objects/AccelStruct.h:61-86) */

/* Threading contract.  The reference object is used from one thread (the game's Lua thread, source/VisTrace.cpp:831-836).
 * Here: a handle may be used from ONE thread at a time for every call that takes HOST pointers (they share per-handle
 * staging buffers) and for populate / refit; calls with VT_TRAVERSE_DEVICE_PTRS may be issued concurrently from several
 * threads when each uses its own stream.  Different handles are independent. */

/* Number of CUDA devices visible; <0 on error. */
int vt_device_count(void);

/* AccelStruct::AccelStruct (source/objects/AccelStruct.cpp:510-523) bound to one GPU.
 * NULL on failure (no CUDA device: the engine has no CPU fallback). */
vt_accel *vt_accel_create(int device);

/* AccelStruct::~AccelStruct (source/objects/AccelStruct.cpp:525-531). */
void vt_accel_destroy(vt_accel *accel);

/* Headless AccelStruct::PopulateAccel (source/objects/AccelStruct.cpp:533-776):
 * copy the scene, derive e1/e2/n/nNorm/lod per triangle (Primitives.h:75-102),
 * build the hierarchy on the host (the step at AccelStruct.cpp:762-770), flatten
 * it to the device layout and upload.  Rebuild = call again. */
int vt_accel_populate(vt_accel *accel, const vt_scene *scene);

/* AccelStruct:Rebuild (source/VisTrace.cpp:798-818) for MOVED geometry of unchanged topology — the same triangles in
 * the same order with new vertices, e.g. props that moved: keeps the structure of the resident hierarchy and refits
 * its boxes bottom-up as bvh::HierarchyRefitter does (libs/bvh/include/bvh/hierarchy_refitter.hpp:20-31, leaf update of
 * libs/bvh/test/refit_bvh.cpp:79-89) instead of the full rebuild the reference performs, then re-derives the resident
 * node layout and uploads.  Fails when nothing was populated or scene->n_tris / n_materials / n_entities differ; materials,
 * entities and textures are taken to be unchanged.  Synchronous; no traversal of this handle may be in flight.  With the
 * default (quad) layout the work runs on the device (K5, vt_refit.cu): the vt_tri_in array is uploaded, the Triangle
 * constructor and the bottom-up requantisation run as kernels; VT_REFIT_DEVICE=0 forces the host path. */
int vt_accel_refit(vt_accel *accel, const vt_scene *scene);

/* The same for ONE contiguous range of the triangle array — the triangles [first, first + count) of the populated scene get
 * the vertices and attributes of tris[0 .. count) (an entity that moved: its triangles are contiguous after ingestion,
 * source/objects/AccelStruct.cpp:561-760); only those records are uploaded.  Needs the resident quad layout.  On failure
 * after the upload (a box left the float grid) the handle is invalid until vt_accel_refit / vt_accel_populate. */
int vt_accel_refit_range(vt_accel *accel, const vt_tri_in *tris, uint64_t first, uint64_t count);

/* Refit quality and the rebuild trigger.  A refit keeps the topology that was built for the ORIGINAL geometry; the further things
 * move, the looser its boxes.  *area_ratio = (sum of the surface areas of all node boxes now) / (the same sum right after the
 * build): the SAH's inner-node term, to which the expected number of node visits is proportional
 * (libs/bvh/include/bvh/sah_based_algorithm.hpp:16-41); 1.0 until the first device-side refit.  vt_accel_set_refit_rebuild_ratio(r),
 * r > 0 (or VT_REFIT_REBUILD_RATIO in the environment): vt_accel_refit rebuilds the hierarchy from scratch — what accel:Rebuild
 * always does in the reference (source/VisTrace.cpp:798-818) — whenever the refitted tree exceeds r; *rebuilds counts those.
 * vt_accel_refit_range cannot rebuild (it does not see the whole scene): the caller polls the ratio and calls vt_accel_populate. */
int vt_accel_refit_quality(const vt_accel *accel, double *area_ratio, uint64_t *rebuilds);
int vt_accel_set_refit_rebuild_ratio(vt_accel *accel, double ratio);

/* Host-only: the refit step alone — `nodes` (bvh::Bvh<float> form, node_count entries) are updated in place for the
 * triangles of `scene`; prim_indices has scene->n_tris entries. */
int vt_refit_bvh(const vt_scene *scene, vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices);

/* Same, but with a hierarchy the caller already built, in bvh::Bvh<float> form
 * (`mAccel.nodes`, `mAccel.primitive_indices`, `mAccel.node_count`;
 * libs/bvh/include/bvh/bvh.hpp:96-99).  This is what the reference's own
 * PopulateAccel would pass right after LeafCollapser::collapse
 * (source/objects/AccelStruct.cpp:769-770). */
int vt_accel_populate_with_bvh(vt_accel *accel, const vt_scene *scene,
                               const vt_node *nodes, uint64_t node_count,
                               const uint64_t *prim_indices);

/* Read the hierarchy back in bvh::Bvh<float> form.  Call with nodes == NULL to get
 * the counts.  prim_indices has n_tris entries. */
int vt_accel_get_bvh(const vt_accel *accel, vt_node *nodes, uint64_t *node_count,
                     uint64_t *prim_indices, uint64_t *n_tris);

/* Batched AccelStruct::Traverse (source/objects/AccelStruct.cpp:778-838).
 *   rays/hits/attrs : n records each; host pointers unless VT_TRAVERSE_DEVICE_PTRS.
 *   attrs           : NULL = hit record only; else the eager TraceResult per ray.
 *   stream          : cudaStream_t or NULL (legacy default stream).
 * Host pointers: synchronous — results are in `hits` on return.  Device pointers:
 * enqueued on `stream`, returns immediately.
 * Per-ray argument rules of the reference (tMin < 0, tMax <= tMin: AccelStruct.cpp:805-806)
 * turn the ray into a miss and are counted in vt_accel_invalid_rays(). */
#define VT_TRAVERSE_DEVICE_PTRS 1u
#define VT_TRAVERSE_ANY_HIT     2u /* early-out occlusion query: hit.prim != VT_MISS is all that is defined */
int vt_accel_traverse(vt_accel *accel, const vt_ray *rays, uint64_t n, vt_hit *hits,
                      vt_attr *attrs, uint32_t flags, void *stream);

/* The opt-in statistics overload of the reference traverser (SingleRayTraverser::Statistics,
 * libs/bvh/include/bvh/single_ray_traverser.hpp:132-135,158-163) summed over a batch of closest-hit
 * queries: *steps = traversal steps (sibling-pair visits), *tests = primitive intersections.
 * Synchronous; rays are host pointers unless VT_TRAVERSE_DEVICE_PTRS.  On the exact layout the totals
 * equal the reference's; on the compact layout steps is slightly larger (conservative boxes). */
int vt_accel_traverse_stats(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t flags,
                            uint64_t *steps, uint64_t *tests);
/* The same per ray (quad / compact layouts): per_ray[i] = steps | tests << 16 of ray i, each saturating at 65535 — the length of
 * the ray's dependent chain, which bounds the duration of small launches.  per_ray is a HOST array of n words. */
int vt_accel_traverse_ray_stats(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t flags, uint32_t *per_ray);

/* Same with per-ray texture-LOD cones: cones = n x {coneWidth, coneAngle}, the 5th and 6th
 * arguments of accel:Traverse (source/objects/AccelStruct.cpp:795-803).  NULL = (-1, -1) = mip 0. */
int vt_accel_traverse_cones(vt_accel *accel, const vt_ray *rays, const float *cones, uint64_t n,
                            vt_hit *hits, vt_attr *attrs, uint32_t flags, void *stream);

/* Eager TraceResult for hits that were produced earlier (device or host pointers as
 * per flags): the constructor + getters of source/objects/TraceResult.cpp:45-262. */
int vt_accel_trace_result(vt_accel *accel, const vt_ray *rays, const vt_hit *hits,
                          uint64_t n, vt_attr *attrs, uint32_t flags, void *stream);

/* Secondary-ray generation on the device, the step GLua scripts do per hit between two
 * accel:Traverse calls: for every non-sky hit in attrs[0, n) spawn `spp` cosine-weighted bounce
 * rays about the shading normal — hemisphere_cos (source/libraries/BSDF.cpp:69-77) in the hit's
 * tangent/binormal/normal frame, origin = vistrace.CalcRayOrigin(pos, geometric normal)
 * (source/VisTrace.cpp:1478-1519) — into out_rays[i*spp + s].  Slots that spawn nothing (miss,
 * sky) are MASKED (tmax < 0): vt_accel_traverse reports them as misses and does not count them.
 * Random numbers are a counter-based hash of (slot, dimension, seed).  live_out (nullable, host
 * pointer) receives the number of rays spawned and makes the call synchronous. */
int vt_accel_bounce_rays(vt_accel *accel, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed,
                         vt_ray *out_rays, uint64_t *live_out, uint32_t flags, void *stream);

/* Batched SampleBSDF, DIFFUSE LOBE (source/libraries/BSDF.cpp:770-825, SampleDiffuse :252-278, hemisphere_cos :69-77) — the
 * other call a GLua path tracer makes per hit between two accel:Traverse calls.  For every non-sky hit in attrs[0, n) and every
 * sample s < spp: a BSDFMaterial whose activeLobes is LobeType::DiffuseReflection (the remaining fields at their defaults,
 * BSDF.h:58-92) is prepared by PrepShadingData(albedo, metalness, roughness) (BSDF.cpp:11-21) from the TraceResult record;
 * incident = TraceResult's wo = -normalize(direction of rays[i]) (source/objects/AccelStruct.cpp:826); the three numbers the
 * reference draws from its ISampler — lobeSelect, then r1, r2 of hemisphere_cos — are vt_sample_uniform01(i*spp+s, 0 / 1 / 2, seed).
 * out_samples[i*spp+s] = the BSDFSample (scattered in world space, pdf, weight, lobe); out_rays[i*spp+s] = the ray along
 * `scattered` from vistrace.CalcRayOrigin(pos, geometric normal on the side `scattered` leaves through)
 * (source/VisTrace.cpp:1478-1519), tMax = FLT_MAX.  Slots that spawn nothing — miss, sky, or lobeSelect >= pDiffuse (a fully
 * metallic hit: SampleBSDF returns with the zero vector, weight 0, pdf 0, lobe None) — are MASKED rays (tmax < 0).
 * The specular / transmission lobes are not restated (SURVEY.md section 8 f2 names the diffuse lobe).
 * Host pointers unless VT_TRAVERSE_DEVICE_PTRS; live_out (nullable, host) = rays spawned, makes the call synchronous;
 * queue / queue_count / miss_hits (all or none; device pointers only) as in vt_accel_bounce_rays_queued. */
int vt_accel_sample_bsdf_rays(vt_accel *accel, const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed,
                              vt_ray *out_rays, vt_bsdf_sample *out_samples, uint64_t *live_out, uint32_t *queue,
                              uint64_t *queue_count, vt_hit *miss_hits, uint32_t flags, void *stream);

/* Host-only: the counter-based random number the generators draw for (slot, dimension, seed), in [0, 1) — the stand-in for the
 * reference's sequential mt19937 Sampler (source/objects/Sampler.cpp:5-20), exposed so that a caller can reproduce a wave. */
float vt_sample_uniform01(uint64_t slot, uint32_t dim, uint64_t seed);

/* Shadow-ray generation on the device ("primary + shadow rays"): for every non-sky hit in attrs[0, n) one ray from
 * vistrace.CalcRayOrigin(pos, geometric normal on the viewer's side) (source/VisTrace.cpp:1478-1519) into out_rays[i]:
 *   point_light == 0: direction = light (a sun direction, used as given), tMax = tmax;
 *   point_light != 0: direction = light - origin, tMax = 1 (t is parametric, source/objects/AccelStruct.cpp:810-815).
 * Misses and sky hits leave MASKED slots (tmax < 0).  Trace the result with VT_TRAVERSE_ANY_HIT for an occlusion
 * query.  live_out (nullable, host pointer) receives the number of rays spawned and makes the call synchronous. */
int vt_accel_shadow_rays(vt_accel *accel, const vt_attr *attrs, uint64_t n, const float light[3], int point_light,
                         float tmax, vt_ray *out_rays, uint64_t *live_out, uint32_t flags, void *stream);

/* Ray queue — wavefront compaction between two accel:Traverse waves.  A generator call with a queue also LISTS the
 * slots it filled: queue[0, *queue_count) (u32 slot indices; *queue_count is zeroed by the call and counted on the
 * device) and writes the miss record of every slot it masked into miss_hits[slot].  vt_accel_traverse_queued then
 * traces rays[queue[k]] -> hits[queue[k]] for k < min(capacity, *queue_count) and touches nothing else, so masked
 * slots never occupy a lane, yet hits[0, capacity) is complete in slot order when miss_hits == hits (required when
 * attrs are requested: the TraceResult stage reads every slot's hit record).  Same per-ray
 * semantics as vt_accel_traverse (source/objects/AccelStruct.cpp:778-838); the order of the queue is unspecified, the
 * hit buffer does not depend on it.  DEVICE pointers only, everything is enqueued on `stream`; at most 2^32 slots. */
int vt_accel_bounce_rays_queued(vt_accel *accel, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed,
                                vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream);
int vt_accel_shadow_rays_queued(vt_accel *accel, const vt_attr *attrs, uint64_t n, const float light[3], int point_light,
                                float tmax, vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits,
                                void *stream);
int vt_accel_traverse_queued(vt_accel *accel, const vt_ray *rays, const uint32_t *queue, const uint64_t *queue_count,
                             uint64_t capacity, vt_hit *hits, vt_attr *attrs /* nullable: TraceResult of all `capacity` slots */,
                             uint32_t flags /* VT_TRAVERSE_ANY_HIT */, void *stream);

/* Wave compaction for multi-bounce paths (path tracing: primary, then bounce after bounce, each with its shadow rays).  After
 * the first bounce most slots of a wave are dead — the path left through the sky — and a wave that still visits every slot
 * spends its time on them.  The requeued generators visit only the parents the PREVIOUS wave's queue lists
 * (attrs[in_queue[k]], k < min(n, *in_count)): slot = parent * spp + sample as before, slots they do not visit are not
 * written at all.  With VT_TRAVERSE_QUEUE_ATTRS vt_accel_traverse_queued likewise builds the TraceResult only of the slots
 * its queue lists.  Everything downstream of a compacted wave must therefore go through the queues; the per-ray results are
 * those of the uncompacted calls (same rays, same random-number counters, same hits).  in_queue must differ from queue.
 * DEVICE pointers, enqueued on `stream`. */
#define VT_TRAVERSE_QUEUE_ATTRS 4u
int vt_accel_bounce_rays_requeued(vt_accel *accel, const vt_attr *attrs, const uint32_t *in_queue, const uint64_t *in_count,
                                  uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays, uint32_t *queue,
                                  uint64_t *queue_count, vt_hit *miss_hits, void *stream);
int vt_accel_shadow_rays_requeued(vt_accel *accel, const vt_attr *attrs, const uint32_t *in_queue, const uint64_t *in_count,
                                  uint64_t n, const float light[3], int point_light, float tmax, vt_ray *out_rays,
                                  uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream);

/* Path waves in one call (BASELINE.json configs[4]: "per sample primary + shadow + up to K diffuse bounces each with a shadow ray"):
 * rays[0, n) are the primary rays of n paths (slot = pixel).  Wave 0 = accel:Traverse + TraceResult of every ray; wave k + 1 =
 * one cosine-weighted bounce ray (vt_accel_bounce_rays) per vertex of wave k that is still alive (hit something that is not the
 * sky), traversed closest-hit + TraceResult; at every vertex one shadow ray toward sun_dir, traversed any-hit.  From wave 1 on
 * every kernel walks the queue of live paths only (requeued generators, VT_TRAVERSE_QUEUE_ATTRS) unless VT_PATHS_NO_COMPACTION.
 * Harness-level shading folds the result into the RGBFFF framebuffer: fb[i] += weight * (sum over vertices of throughput * sun_rgb
 * where the shadow ray is unoccluded  +  throughput * albedo where the path leaves through a sky brush), throughput = product of
 * the albedos along the path.  ray_counts (nullable, HOST, 2 + 2 * bounces entries; makes the call synchronous): rays traced per
 * wave — [0] primary, [1] its shadow rays, [2 + 2k], [3 + 2k] bounce k + 1 and its shadow rays.
 * DEVICE pointers (VT_TRAVERSE_DEVICE_PTRS required), enqueued on `stream`; at most 8 bounces. */
#define VT_PATHS_NO_COMPACTION 8u
/* the call uses the handle's SECOND set of path scratch buffers: calls with and without the flag may run concurrently on two streams
 * (two samples of a frame in flight: one sample's small late waves run under the other's large early ones) — into DIFFERENT
 * framebuffers, which the caller adds up */
#define VT_PATHS_SLOT1 16u
int vt_accel_trace_paths(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t bounces, const float sun_dir[3],
                         const float sun_rgb[3], uint64_t seed, float weight, float *framebuffer_rgb, uint64_t *ray_counts,
                         uint32_t flags, void *stream);

/* The "primary + diffuse" wave of the headline benchmark in one call: traverse rays[0, n), build
 * the TraceResult of every hit, spawn spp bounce rays per hit (as vt_accel_bounce_rays) and
 * traverse those.  Out: hits[n], bounce_hits[n*spp]; optional attrs[n], bounce_rays[n*spp]
 * (required scratch with VT_TRAVERSE_DEVICE_PTRS).  Host pointers are processed in tiles over
 * several CUDA streams so the host<->device copies overlap the kernels; the call returns when
 * all results are on the host.  Device pointers: enqueued on `stream`, live_out must be NULL. */
int vt_accel_trace_diffuse_wave(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed,
                                vt_hit *hits, vt_attr *attrs, vt_ray *bounce_rays, vt_hit *bounce_hits,
                                uint64_t *live_out, uint32_t flags, void *stream);

/* Fold one diffuse wave into an RGBFFF framebuffer — tightly packed 3 x f32 per pixel, the memory
 * an IRenderTarget of format RGBFFF exposes through GetRawData(0) (include/vistrace/IRenderTarget.h:40):
 *   fb[i] += weight * albedo_i * (bounce rays of pixel i that escape to the sky) / spp.
 * DEVICE pointers only; enqueued on `stream`.  This per-GPU partial image is what the multi-GPU path
 * sums with its single NCCL collective. */
int vt_accel_accumulate_sky(vt_accel *accel, const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n,
                            uint32_t spp, float weight, float *framebuffer_rgb, void *stream);

/* Node layout of the device-resident hierarchy, applied by the next populate call.
 *   VT_LAYOUT_COMPACT (default): each sibling pair — the two Bvh::Node records one traversal step
 *     reads (libs/bvh/include/bvh/single_ray_traverser.hpp:85-87) — is stored in 32 bytes with
 *     conservatively quantised boxes.  Every node FastNodeIntersector (node_intersectors.hpp:35-47)
 *     accepts is still visited, the triangle test is the exact one, so t/u/v/prim are bit-identical to
 *     the reference whenever the closest hit is unique; only the winner among candidates whose t is
 *     equal (or differs by float rounding) can differ, because equal-distance subtrees may be
 *     visited in another order.  Trees it cannot hold (a leaf of > 15 triangles, non-finite bounds)
 *     fall back to the exact layout: query vt_accel_get_layout() after populate.
 *   VT_LAYOUT_EXACT: the 2 x 32-byte nodes verbatim; visit order and exact-tie winners are the
 *     reference's (single_ray_traverser.hpp:55-60,109-115). */
#define VT_LAYOUT_EXACT   0
#define VT_LAYOUT_COMPACT 1
#define VT_LAYOUT_QUAD    2
int vt_accel_set_layout(vt_accel *accel, int layout);
int vt_accel_get_layout(const vt_accel *accel);

/* The diffuse wave with the framebuffer as its only result: HOST rays[n] in, HOST framebuffer_rgb[3n] out —
 * the memory of an RGBFFF IRenderTarget (include/vistrace/IRenderTarget.h:40; GetRawData(0)), written, not
 * accumulated: fb[i] = weight * albedo_i * (escaped fraction of pixel i's spp bounce rays), sky pixels =
 * weight * albedo.  Runs vt_accel_trace_diffuse_wave + vt_accel_accumulate_sky per tile on several CUDA streams,
 * so the ray upload, the kernels and the image download overlap; returns when the image is on the host.
 * live_out (nullable) receives the number of bounce rays spawned. */
int vt_accel_render_diffuse_wave(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed,
                                 float weight, float *framebuffer_rgb, uint64_t *live_out);
/* The same call split in two, so that a host can keep TWO frames in flight: _begin enqueues the frame (uploads, kernels, download) and
 * returns; _wait blocks until the OLDEST frame begun on this handle is complete in its framebuffer (frames complete in the order they
 * were begun; at most two may be outstanding).  Frame k's last tiles, launch tails and download then run under frame k + 1's first
 * uploads and kernels.  rays and framebuffer_rgb must stay valid (and pinned, for the copies to be asynchronous) until the frame's
 * _wait returns; consecutive frames need different framebuffers. */
int vt_accel_render_diffuse_wave_begin(vt_accel *accel, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight,
                                       float *framebuffer_rgb);
int vt_accel_render_diffuse_wave_wait(vt_accel *accel);

/* Rays rejected by the argument rules during the last synchronous traverse. */
uint64_t vt_accel_invalid_rays(const vt_accel *accel);

/* Kernel launches issued by this handle so far (bench.py's gpu_launches). */
uint64_t vt_accel_launch_count(const vt_accel *accel);

/* Scene statistics: n_tris, node_count, bytes resident in HBM. */
int vt_accel_stats(const vt_accel *accel, uint64_t *n_tris, uint64_t *node_count,
                   uint64_t *device_bytes);

/* What the Triangle constructor derived per input triangle, n x 16 floats
 * {p0, e1, e2, n, nNorm, lod} (source/objects/Primitives.h:75-102). */
int vt_accel_get_tri_derived(const vt_accel *accel, float *out16);

/* Host-only (no GPU touched): the hierarchy build step on its own, bvh::Bvh<float> form out.
 * Call with nodes == NULL to get the node count; *node_count is the capacity of `nodes` on
 * entry and the node count on return; prim_indices has scene->n_tris entries. */
int vt_build_bvh(const vt_scene *scene, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices);

/* Host-only builder-quality option: reinsertion optimisation of a finished bvh::Bvh<float>-form hierarchy,
 * in place — this library's version of bvh::ParallelReinsertionOptimizer (libs/bvh/include/bvh/parallel_reinsertion_optimizer.hpp),
 * which the reference's library offers and the reference does not call.  `iterations` passes; each takes up to ~`fraction` of the
 * nodes (largest boxes first; 0.05 is a good value) out of the tree and puts them back where the sum of the inner-node areas (the
 * SAH traversal term) grows least.  Leaves keep their primitive ranges, so prim_indices stays valid; node order is depth-first
 * again afterwards.  The array is left untouched when the result would be deeper than 60 levels or no better.  area_before /
 * area_after (nullable): the sum of inner-node half-areas; moves (nullable): reinsertions applied.  VT_REINSERT=<iterations> makes
 * vt_accel_populate / vt_build_bvh run it on the product builder's tree (off by default: profiles/r2_child_order.md section 3). */
int vt_optimize_bvh(vt_node *nodes, uint64_t node_count, int iterations, double fraction, double *area_before, double *area_after,
                    uint64_t *moves);

/* Host-only: the REFERENCE's hierarchy for the scene, rebuilt from its algorithm — bvh::LocallyOrderedClusteringBuilder
 * <BVH, uint32_t> (libs/bvh/include/bvh/locally_ordered_clustering_builder.hpp: search radius 14, 30-bit Morton codes) and,
 * with collapse != 0, bvh::LeafCollapser (leaf_collapser.hpp) — the sequence of source/objects/AccelStruct.cpp:762-770.
 * Nodes and primitive indices equal the reference's arrays bit for bit, so the exact node layout then reproduces the
 * reference's winners among exactly tied candidates.  Same calling convention as vt_build_bvh; VT_BUILDER=ploc makes
 * vt_accel_populate use it. */
int vt_build_bvh_ploc(const vt_scene *scene, int collapse, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices);

/* Host-only: flatten a bvh::Bvh<float>-form hierarchy to the device layout — (node_count-1)/2
 * 64-byte sibling pairs (the first bfs_pairs in breadth-first order, depth-first below) and the
 * leaf-order permutation of the triangles.  Validates the tree (adjacent odd child pairs, every
 * primitive covered once, depth <= 64). */
int vt_flatten_bvh(const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices, uint64_t n_tris,
                   uint32_t bfs_pairs, void *pairs_out, uint32_t *leaf_order_out,
                   uint32_t *root_leaf_count, uint32_t *max_depth);

/* Host-only: SkinTriangle (source/objects/AccelStruct.cpp:66-108) over n triangles, in place — what
 * PopulateAccel runs per entity triangle before the build (AccelStruct.cpp:742-749): positions
 * {p0, p0 - e1, p0 + e2} (e1 = p0 - p1, e2 = p2 - p0 as rounded by the Triangle constructor), normals and
 * tangents become sum_i bones[b_i] * binds[b_i] * vec4(v, w) * weight_i (w = 1 for positions, 0 for
 * directions; glm::mat4 arithmetic order), then p = the new vertices; e1/e2/n/nNorm/lod are re-derived by
 * vt_accel_populate.  bones/binds: n_bones glm::mat4 each, 16 floats, column-major.  skin == NULL applies
 * the one-bone overload (AccelStruct.cpp:103-108): every vertex bound to bone 0 with weight 1. */
int vt_skin_triangles(vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones,
                      const float *binds, uint32_t n_bones);

/* Host-only: 64-byte sibling pairs in depth-first order (vt_flatten_bvh with bfs_pairs = 0) ->
 * n_pairs 32-byte compact pairs: per axis {origin_adj f32} x3, {biased exponent u8} x3,
 * counts u8 (lcount | rcount << 4), q[axis][l.lo, l.hi, r.lo, r.hi] u8, ref u32.
 * plane = (2^23 + q) * 2^E + origin_adj, exactly, with lo' <= lo and hi' >= hi. */
int vt_compact_pairs(const void *pairs, uint64_t n_pairs, void *cpairs_out);

/* Host-only: bvh::Bvh<float>-form hierarchy -> quad nodes (64 bytes each, depth-first order; layout in
 * vistrace_b200/csrc/vt_device.h: VtQuad) + the leaf-order permutation of the triangles.  Call with
 * quads_out == NULL to get the count; *n_quads is the capacity on entry, the count on return.
 * *max_stack = worst-case number of pending child references during traversal (<= 64).
 * plane = (OFFSET + q) * scale + origin_adj exactly, OFFSET = vt_quad_plane_offset() (the float the kernel
 * forms from a quantised byte: 1024 through the fp16 pattern 0x6400 | q, or 2^23 through 0x4B000000 | q). */
uint32_t vt_quad_plane_offset(void);
int vt_build_quads(const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices, uint64_t n_tris,
                   void *quads_out, uint64_t *n_quads, uint32_t *leaf_order_out, uint32_t *root_leaf_count,
                   uint32_t *max_stack);

/* ------------------------------------------------------------------ multi-GPU
 * The reference is single-threaded CPU code with no multi-device notion; the north star shards ray
 * batches over the GPUs of one box behind the same object: the hierarchy is built ONCE, its device image is replicated on
 * every GPU (the scene is read-only and far smaller than 180 GB), rays are partitioned, and the only communication is the
 * final gather of hit-buffer slices or framebuffer tiles.  vt_group stands for "one AccelStruct
 * (source/objects/AccelStruct.h:61-86) resident on several GPUs".
 *
 *   vt_group_create(devices, n)        one process drives n local GPUs (a worker thread per GPU; every GPU reads its share of
 *                                      the caller's host buffers and writes its results straight back over its own PCIe link;
 *                                      replication by cudaMemcpyPeer).
 *   vt_group_create_rank(dev, r, w, id)  one process per GPU (torchrun / MPI style): `id` is the 128-byte ncclUniqueId that
 *                                      rank 0 obtained from vt_group_unique_id() and the launcher distributed.  Rank 0 builds,
 *                                      the image is ncclBroadcast, results are gathered on rank 0 over NVLink (ncclSend / ncclRecv).
 *                                      NCCL is bound at run time (libnccl.so.2 via dlopen; VT_NCCL_LIB overrides the name).
 * Every call on a multi-process group is COLLECTIVE: all ranks call it with the same n / spp / seed / flags. */
typedef struct vt_group vt_group;
int vt_group_unique_id(uint8_t id[128]);
vt_group *vt_group_create(const int *devices, int n);
vt_group *vt_group_create_rank(int device, int rank, int world, const uint8_t id[128]);
void vt_group_destroy(vt_group *group);
int vt_group_size(const vt_group *group);          /* GPUs in the group (world size) */
int vt_group_rank(const vt_group *group);          /* global rank of this process's first member */
int vt_group_local_members(const vt_group *group); /* members driven by this process */
vt_accel *vt_group_accel(vt_group *group, int local_member); /* borrowed; members other than the builder are replicas (no refit / get_bvh) */

/* AccelStruct::PopulateAccel (source/objects/AccelStruct.cpp:533-776) for the whole group: ingest + build + flatten once,
 * replicate the device image.  Multi-process groups: only rank 0 reads `scene` (the others may pass NULL). */
int vt_group_populate(vt_group *group, const vt_scene *scene);

/* Batched AccelStruct::Traverse (source/objects/AccelStruct.cpp:778-838) over the group: contiguous, balanced slices of
 * rays[0, n) — rank r traces [r * n / w ...) — HOST pointers.  Single process: hits / attrs are complete on return.
 * Multi-process: every rank passes frame-sized arrays and reads only its own slice of `rays`; the slices are gathered on
 * rank 0 over NVLink and downloaded there (rank 0: complete arrays; other ranks: their own slice). */
int vt_group_traverse(vt_group *group, const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, uint32_t flags);

/* Frame sharding used by vt_group_render_diffuse_wave: the frame of n pixels is cut into tiles of *tile pixels dealt
 * round-robin to the ranks (tile g belongs to rank g % w); *local_count = pixels of `rank`, stored compactly in tile order. */
int vt_group_shard(const vt_group *group, uint64_t n, int rank, uint64_t *tile, uint64_t *local_count);

/* Host-only (no GPU touched): the same geometry for any (world, rank, tile) — what a launcher needs to lay out per-rank buffers. */
int vt_shard_geometry(uint64_t n, int world, int rank, uint64_t tile, uint64_t *local_count, uint64_t *local_tiles);

/* vt_accel_render_diffuse_wave over the group, STRONG scaling of one frame: every pixel is traced exactly once, by the rank
 * that owns its tile, and the image equals the single-GPU image bit for bit (the bounce rays' random-number counters come
 * from global pixel indices).  flags = 0: HOST rays[n] in, HOST RGBFFF framebuffer_rgb[3n] out (include/vistrace/IRenderTarget.h:40),
 * complete on return in a single-process group and on rank 0 of a multi-process group (the other ranks' framebuffer is not
 * written: their GPUs store finished pixels straight into rank 0's frame over NVLink — peer memory, CUDA IPC — which rank 0
 * downloads stretch by stretch as the ranks deliver; VT_GROUP_GATHER=nccl: ncclSend / ncclRecv of the shards instead);
 * *live_out (nullable) = bounce rays spawned by THIS process's GPUs.
 * flags = VT_TRAVERSE_DEVICE_PTRS (multi-process groups): rays = this rank's COMPACT shard already resident on its GPU
 * (vt_group_shard: local_count records in tile order), framebuffer_rgb = frame-sized DEVICE image, complete on rank 0; everything
 * is enqueued on `stream` and the call returns without synchronising; live_out must be NULL.
 * flags = VT_GROUP_SHARED_HOST_FRAME (multi-process groups on one node, host pointers): framebuffer_rgb is the SAME host memory in
 * every process — a shared mapping (POSIX shm) that each process pinned (cudaHostRegister) — e.g. the RGBFFF render target of the
 * process that displays it.  Every rank lands its own tiles through its own PCIe link; nothing crosses NVLink, nothing funnels through
 * rank 0's link; a one-byte ncclAllGather behind the copies makes the frame complete on EVERY rank's return. */
#define VT_GROUP_SHARED_HOST_FRAME 16u
/* With VT_TRAVERSE_DEVICE_PTRS (multi-process groups): the call uses the group's SECOND set of per-frame state — its own peer-memory
 * frame and hand-shake flags on rank 0, its own scratch buffers.  Calls with and without the flag share nothing, so a caller that
 * alternates consecutive frames between two streams (frame k: stream A, flags without; frame k + 1: stream B, with) has two frames
 * in flight: one frame's launch tails run under the other frame's bulk (profiles/r2_tail_sharing.md: 1.4x on a 1/8-frame shard).
 * Calls on the SAME slot must be issued in stream order on one stream, in the same order on every rank. */
#define VT_GROUP_FRAME_SLOT1 32u
/* With VT_GROUP_SHARED_HOST_FRAME: the call enqueues the frame and returns; vt_group_wait_frame blocks until the OLDEST such frame is
 * complete in its shared framebuffer on every rank (at most two in flight; consecutive frames need different framebuffers; rays and
 * framebuffer stay valid until the frame's wait returns; live_out must be NULL).  Two frames in flight hide one frame's last tiles,
 * launch tails and download under the next frame's uploads and kernels. */
#define VT_GROUP_ASYNC 64u
int vt_group_render_diffuse_wave(vt_group *group, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight,
                                 float *framebuffer_rgb, uint64_t *live_out, uint32_t flags, void *stream);
int vt_group_wait_frame(vt_group *group); /* collective like the call that began the frame */

/* Sample-index sharding (every rank renders the whole frame for its own samples): sum the per-rank DEVICE images into
 * rank 0's, in place — one ncclReduce over NVLink, enqueued on `stream`.  Multi-process groups. */
int vt_group_reduce_device(vt_group *group, float *buf, uint64_t count, void *stream);

/* The other collective of a sample-index-sharded frame: every rank has uploaded ITS 1 / w of a device buffer (bytes_per_rank
 * bytes at buf + rank * bytes_per_rank, e.g. its slice of the primary rays) and receives the other slices over NVLink —
 * one ncclAllGather, in place, enqueued on `stream`.  Multi-process groups. */
int vt_group_all_gather_device(vt_group *group, void *buf, uint64_t bytes_per_rank, void *stream);

/* Kernel launches + collectives issued through the group so far. */
uint64_t vt_group_launch_count(const vt_group *group);

/* Last error message of the calling thread ("" if none). */
const char *vt_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VISTRACE_B200_H */
