//! Drop-in replacement for the render loop of the reference's `main` (src/main.zig:378-402).
//! UNCOMPILED IN THIS REPOSITORY (no Zig toolchain in the build image).
//!
//! In src/main.zig, everything up to and including `Camera.init(...)` (line 376) and the
//! `Image.create` / `image.writeToFilePath` calls stay as they are; the triple loop between them becomes:
//!
//!     const render_cuda = @import("render_cuda.zig");
//!     try render_cuda.render(allocator, world, camera, background, image_width, image_height,
//!                            samples_per_pixel, max_depth, std.mem.sliceAsBytes(image.pixels.rgb24));
//!
//! `Camera` must be made `pub` (or this function moved into main.zig): it reads the ten fields of
//! src/main.zig:40-51.  The rgb24 pixel array has exactly the byte layout rtw_cuda_render writes
//! (3 bytes per pixel, row 0 = top: the library applies the (H-1-j) flip of main.zig:396 itself).
const std = @import("std");
const rtw = @import("rtw.zig");
const abi = @import("rtw_cuda.zig");
const flatten = @import("flatten.zig").flatten;

fn v3(v: rtw.vec.Vec3) [3]f64 {
    return .{ v.x, v.y, v.z };
}

pub fn render(
    allocator: std.mem.Allocator,
    world: rtw.hittable.Hittable,
    camera: anytype, // the reference's private `Camera` struct (src/main.zig:40-51)
    background: rtw.vec.Color,
    image_width: u32,
    image_height: u32,
    samples_per_pixel: u32,
    max_depth: u32,
    rgb24_out: []u8,
) !void {
    std.debug.assert(rgb24_out.len == @as(usize, image_width) * image_height * 3);
    var flat = try flatten(allocator, world, camera.time0, camera.time1);
    defer flat.deinit();
    const desc = flat.desc();

    var ctx: ?*abi.Ctx = null;
    try abi.check(abi.rtw_cuda_create(0, &ctx), null, error.CudaUnavailable);
    defer abi.rtw_cuda_destroy(ctx);
    try abi.check(abi.rtw_cuda_upload_scene(ctx.?, &desc), ctx, error.SceneRejected);

    const cam = abi.Camera{
        .origin = v3(camera.origin),
        .horizontal = v3(camera.horizontal),
        .vertical = v3(camera.vertical),
        .lower_left_corner = v3(camera.lower_left_corner),
        .u = v3(camera.u),
        .v = v3(camera.v),
        .w = v3(camera.w),
        .lens_radius = camera.lens_radius,
        .time0 = camera.time0,
        .time1 = camera.time1,
    };
    const params = abi.RenderParams{
        .width = image_width,
        .height = image_height,
        .spp_begin = 0,
        .spp_end = samples_per_pixel,
        .spp_total = samples_per_pixel,
        .max_depth = max_depth,
        .seed = 42, // src/main.zig:300
        .background = v3(background),
    };
    try abi.check(abi.rtw_cuda_render(ctx.?, &cam, &params, rgb24_out.ptr, null), ctx, error.RenderFailed);
}
