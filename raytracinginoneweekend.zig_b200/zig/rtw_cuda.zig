//! Zig declarations of the C ABI in include/rtw_cuda.h (`extern struct` = C layout).
//! UNCOMPILED IN THIS REPOSITORY'S CI: the build image has no Zig toolchain (SURVEY.md §0 D7).
//! Field order and types mirror the header one for one; tests/test_abi.py pins the C side.
const std = @import("std");

pub const ABI_VERSION: u32 = 3;
pub const MISS: u32 = 0xFFFFFFFF;

pub const PrimKind = enum(u32) { sphere = 0, moving_sphere = 1, xy_rect = 2, xz_rect = 3, yz_rect = 4 };
pub const XformKind = enum(u32) { translate = 0, rotate_y = 1 };
pub const MatKind = enum(u32) { diffuse = 0, metal = 1, dielectric = 2, diffuse_light = 3 };
pub const TexKind = enum(u32) { solid = 0, checker = 1, noise = 2, image = 3 };
pub const Variant = enum(u32) { auto = 0, mega_flat = 1, mega_bvh = 2, wavefront = 3 };
pub const FLAG_COUNT_EVENTS: u32 = 1;
pub const FLAG_DETERMINISTIC: u32 = 2;

pub const Prim = extern struct {
    kind: u32,
    material: u32,
    xform: i32,
    reserved: u32 = 0,
    v: [10]f64 = [_]f64{0} ** 10,
};
pub const Xform = extern struct { kind: u32, outer: i32, v: [4]f64 = [_]f64{0} ** 4 };
pub const Material = extern struct { kind: u32, texture: i32 = -1, albedo: [3]f64 = .{ 0, 0, 0 }, param: f64 = 0 };
pub const Texture = extern struct {
    kind: u32,
    a: i32 = -1,
    b: i32 = -1,
    reserved: u32 = 0,
    color: [3]f64 = .{ 0, 0, 0 },
    scale: f64 = 0,
};
pub const Image = extern struct { width: u32, height: u32, rgba8: [*]const u8 };
pub const Perlin = extern struct { ranvec: [*]const f64, perm_x: [*]const u32, perm_y: [*]const u32, perm_z: [*]const u32 };
pub const SceneDesc = extern struct {
    n_prims: u32,
    prims: ?[*]const Prim,
    n_xforms: u32,
    xforms: ?[*]const Xform,
    n_materials: u32,
    materials: ?[*]const Material,
    n_textures: u32,
    textures: ?[*]const Texture,
    n_images: u32,
    images: ?[*]const Image,
    n_perlins: u32,
    perlins: ?[*]const Perlin,
    time0: f64,
    time1: f64,
};
pub const Camera = extern struct {
    origin: [3]f64,
    horizontal: [3]f64,
    vertical: [3]f64,
    lower_left_corner: [3]f64,
    u: [3]f64,
    v: [3]f64,
    w: [3]f64,
    lens_radius: f64,
    time0: f64,
    time1: f64,
};
pub const RenderParams = extern struct {
    width: u32,
    height: u32,
    spp_begin: u32,
    spp_end: u32,
    spp_total: u32,
    max_depth: u32,
    variant: u32 = 0,
    flags: u32 = 0,
    seed: u64 = 42,
    background: [3]f64,
};
pub const Ctx = opaque {};

pub extern "c" fn rtw_cuda_create(device: c_int, out: *?*Ctx) c_int;
pub extern "c" fn rtw_cuda_destroy(ctx: ?*Ctx) void;
pub extern "c" fn rtw_cuda_last_error(ctx: ?*const Ctx) [*:0]const u8;
pub extern "c" fn rtw_cuda_abi_version() u32;
pub extern "c" fn rtw_cuda_upload_scene(ctx: *Ctx, scene: *const SceneDesc) c_int;
pub extern "c" fn rtw_cuda_render(ctx: *Ctx, cam: *const Camera, params: *const RenderParams, rgb8_out: [*]u8, accum_out: ?[*]f32) c_int;

/// rtw_stats of include/rtw_cuda.h: event counters (filled when flags & 1), timers, facts about the uploaded scene.
pub const Stats = extern struct {
    paths: u64, rays: u64, node_tests: u64, sphere_tests: u64, sphere_roots: u64, moving_tests: u64,
    rect_tests: u64, rect_accepts: u64, xform_apps: u64, sphere_finalise: u64, scatter_diffuse: u64,
    scatter_metal: u64, scatter_dielectric: u64, emit_hits: u64, tex_checker: u64, tex_image: u64,
    tex_noise: u64, nan_pixels: u64,
    ms_trace: f64, ms_resolve: f64, ms_upload: f64,
    n_launches: u32, variant_used: u32, bvh_nodes: u32, bvh_depth: u32,
    ms_bvh_build: f64,
    bvh_builder: u32, // 0 = host binned SAH, 1 = device Morton/radix tree
    reserved0: u32,
    ms_wall: f64, // host wall clock of the last render / render_multi call
};
pub extern "c" fn rtw_cuda_stats(ctx: *Ctx, out: *Stats) c_int;

pub extern "c" fn rtw_cuda_create_multi(n_gpus: u32, out: [*]?*Ctx) c_int;
pub extern "c" fn rtw_cuda_set_option(ctx: *Ctx, name: [*:0]const u8, value: ?[*:0]const u8) c_int;
// parity probes of the stochastic device code (include/rtw_cuda.h): results + the random choices they were made from
pub extern "c" fn rtw_cuda_unit_camera(ctx: *Ctx, cam: *const Camera, params: *const RenderParams, n: u32, ijs: [*]const u32, out: [*]f32) c_int;
pub extern "c" fn rtw_cuda_unit_samplers(ctx: *Ctx, n: u32, u3: [*]const f32, out: [*]f32) c_int;
pub extern "c" fn rtw_cuda_unit_uniforms(ctx: *Ctx, params: *const RenderParams, n: u32, psb: [*]const u32, out: [*]f32) c_int;
pub extern "c" fn rtw_cuda_unit_shade(ctx: *Ctx, params: *const RenderParams, n: u32, rays: [*]const f64, psb: [*]const u32, prim_id: [*]u32, out: [*]f32) c_int;
pub extern "c" fn rtw_cuda_render_multi(ctxs: [*]const ?*Ctx, n_ctx: u32, cam: *const Camera, params: *const RenderParams, rgb8_out: [*]u8) c_int;

pub const Error = error{ CudaUnavailable, SceneRejected, RenderFailed };

/// The reference's hot path is infallible (`hit`/`scatter` return bool); only setup can fail, so the
/// integer status codes are folded into a small error set here.
pub fn check(rc: c_int, ctx: ?*const Ctx, comptime e: Error) Error!void {
    if (rc != 0) {
        std.debug.print("rtw_cuda: {s}\n", .{rtw_cuda_last_error(ctx)});
        return e;
    }
}
