//! Scene-flattening pass for the Zig host: walks the reference's ref-counted graph
//! (`rtw.hittable.Hittable`, `rtw.material.Material`, `rtw.texture.Texture`, `Rc(T)`) and emits the
//! plain-old-data arrays of include/rtw_cuda.h.  Meant to live at src/flatten.zig of the reference tree.
//! UNCOMPILED IN THIS REPOSITORY: no Zig toolchain in the build image (SURVEY.md §0 D7).  Its C++ twin
//! (host/rtw_host.cpp, `flatten`) implements the same walk and IS tested (tests/test_host.py).
//!
//! Rules (identical to the C++ twin):
//!  * leaves get ids in depth-first append order — that order is what reproduces the reference's tie rule
//!    ("later list element wins", src/rtw/hittable.zig:235-242);
//!  * `list` vanishes, `box` becomes its 6 side rects (src/rtw/hittable.zig:437-442);
//!  * `translate` / `rotateY` become instance nodes linked OUTWARD (`outer`), shared by the leaves below them;
//!  * materials are deduplicated on the address of the Rc cell payload (`Rc.get()`), i.e. on material identity.
const std = @import("std");
const rtw = @import("rtw.zig");
const abi = @import("rtw_cuda.zig");

const Hittable = rtw.hittable.Hittable;
const Material = rtw.material.Material;
const Texture = rtw.texture.Texture;

pub const FlatScene = struct {
    prims: std.ArrayList(abi.Prim),
    xforms: std.ArrayList(abi.Xform),
    materials: std.ArrayList(abi.Material),
    textures: std.ArrayList(abi.Texture),
    images: std.ArrayList(abi.Image),
    perlins: std.ArrayList(abi.Perlin),
    // storage the POD arrays point into
    ranvec_store: std.ArrayList([]f64),
    perm_store: std.ArrayList([]u32),
    mat_index: std.AutoHashMap(*const Material, u32),
    allocator: std.mem.Allocator,
    time0: f64,
    time1: f64,

    pub fn init(allocator: std.mem.Allocator, time0: f64, time1: f64) FlatScene {
        return .{
            .prims = std.ArrayList(abi.Prim).init(allocator),
            .xforms = std.ArrayList(abi.Xform).init(allocator),
            .materials = std.ArrayList(abi.Material).init(allocator),
            .textures = std.ArrayList(abi.Texture).init(allocator),
            .images = std.ArrayList(abi.Image).init(allocator),
            .perlins = std.ArrayList(abi.Perlin).init(allocator),
            .ranvec_store = std.ArrayList([]f64).init(allocator),
            .perm_store = std.ArrayList([]u32).init(allocator),
            .mat_index = std.AutoHashMap(*const Material, u32).init(allocator),
            .allocator = allocator,
            .time0 = time0,
            .time1 = time1,
        };
    }

    pub fn deinit(self: *FlatScene) void {
        for (self.ranvec_store.items) |s| self.allocator.free(s);
        for (self.perm_store.items) |s| self.allocator.free(s);
        self.ranvec_store.deinit();
        self.perm_store.deinit();
        self.mat_index.deinit();
        self.perlins.deinit();
        self.images.deinit();
        self.textures.deinit();
        self.materials.deinit();
        self.xforms.deinit();
        self.prims.deinit();
    }

    pub fn desc(self: *const FlatScene) abi.SceneDesc {
        return .{
            .n_prims = @intCast(self.prims.items.len),
            .prims = self.prims.items.ptr,
            .n_xforms = @intCast(self.xforms.items.len),
            .xforms = self.xforms.items.ptr,
            .n_materials = @intCast(self.materials.items.len),
            .materials = self.materials.items.ptr,
            .n_textures = @intCast(self.textures.items.len),
            .textures = self.textures.items.ptr,
            .n_images = @intCast(self.images.items.len),
            .images = self.images.items.ptr,
            .n_perlins = @intCast(self.perlins.items.len),
            .perlins = self.perlins.items.ptr,
            .time0 = self.time0,
            .time1 = self.time1,
        };
    }

    fn addTexture(self: *FlatScene, tx: Texture) !i32 {
        var out = abi.Texture{ .kind = 0 };
        switch (tx) {
            .solid => |s| {
                out.kind = @intFromEnum(abi.TexKind.solid);
                out.color = .{ s.color.x, s.color.y, s.color.z };
            },
            .checker => |c| {
                out.kind = @intFromEnum(abi.TexKind.checker);
                out.a = try self.addTexture(c.odd.*);
                out.b = try self.addTexture(c.even.*);
            },
            .noise => |n| {
                out.kind = @intFromEnum(abi.TexKind.noise);
                out.scale = n.scale;
                const rv = try self.allocator.alloc(f64, 256 * 3);
                for (n.perlin.randomVec.items, 0..) |v, i| {
                    rv[3 * i + 0] = v.x;
                    rv[3 * i + 1] = v.y;
                    rv[3 * i + 2] = v.z;
                }
                try self.ranvec_store.append(rv);
                var perms: [3][]u32 = undefined;
                const srcs = [3][]const usize{ n.perlin.permX.items, n.perlin.permY.items, n.perlin.permZ.items };
                for (srcs, 0..) |src, k| {
                    perms[k] = try self.allocator.alloc(u32, 256);
                    for (src, 0..) |p, i| perms[k][i] = @intCast(p);
                    try self.perm_store.append(perms[k]);
                }
                try self.perlins.append(.{ .ranvec = rv.ptr, .perm_x = perms[0].ptr, .perm_y = perms[1].ptr, .perm_z = perms[2].ptr });
                out.a = @intCast(self.perlins.items.len - 1);
            },
            .image => |im| {
                // zigimg hands texture.zig:131-137 4 bytes per texel; the library copies them at upload.
                out.kind = @intFromEnum(abi.TexKind.image);
                try self.images.append(.{
                    .width = @intCast(im.image.width),
                    .height = @intCast(im.image.height),
                    .rgba8 = im.image.pixels.asBytes().ptr,
                });
                out.a = @intCast(self.images.items.len - 1);
            },
        }
        try self.textures.append(out);
        return @intCast(self.textures.items.len - 1);
    }

    fn addMaterial(self: *FlatScene, m: *const Material) !u32 {
        if (self.mat_index.get(m)) |idx| return idx;
        var out = abi.Material{ .kind = 0 };
        switch (m.*) {
            .diffuse => |d| {
                out.kind = @intFromEnum(abi.MatKind.diffuse);
                out.texture = try self.addTexture(d.albedo);
            },
            .metal => |mt| {
                out.kind = @intFromEnum(abi.MatKind.metal);
                out.albedo = .{ mt.albedo.x, mt.albedo.y, mt.albedo.z };
                out.param = mt.fuzz;
            },
            .dielectric => |g| {
                out.kind = @intFromEnum(abi.MatKind.dielectric);
                out.param = g.ir;
            },
            .diffuse_light => |l| {
                out.kind = @intFromEnum(abi.MatKind.diffuse_light);
                out.texture = try self.addTexture(l.emit);
            },
        }
        try self.materials.append(out);
        const idx: u32 = @intCast(self.materials.items.len - 1);
        try self.mat_index.put(m, idx);
        return idx;
    }

    fn addRect(self: *FlatScene, kind: abi.PrimKind, a0: f64, a1: f64, b0: f64, b1: f64, k: f64, m: *const Material, chain: i32) !void {
        var p = abi.Prim{ .kind = @intFromEnum(kind), .material = try self.addMaterial(m), .xform = chain };
        p.v[0] = a0;
        p.v[1] = a1;
        p.v[2] = b0;
        p.v[3] = b1;
        p.v[4] = k;
        try self.prims.append(p);
    }

    /// Depth-first walk; `chain` = index of the innermost instance node enclosing `h`, or -1.
    pub fn walk(self: *FlatScene, h: Hittable, chain: i32) !void {
        switch (h) {
            .sphere => |s| {
                var p = abi.Prim{ .kind = @intFromEnum(abi.PrimKind.sphere), .material = try self.addMaterial(s.material.get()), .xform = chain };
                p.v[0] = s.center.x;
                p.v[1] = s.center.y;
                p.v[2] = s.center.z;
                p.v[3] = s.radius;
                try self.prims.append(p);
            },
            .movingSphere => |s| {
                var p = abi.Prim{ .kind = @intFromEnum(abi.PrimKind.moving_sphere), .material = try self.addMaterial(s.material.get()), .xform = chain };
                p.v = .{ s.center0.x, s.center0.y, s.center0.z, s.center1.x, s.center1.y, s.center1.z, s.time0, s.time1, s.radius, 0 };
                try self.prims.append(p);
            },
            .list => |l| for (l.objects.items) |o| try self.walk(o, chain),
            .xyRect => |r| try self.addRect(.xy_rect, r.x0, r.x1, r.y0, r.y1, r.k, r.material.get(), chain),
            .xzRect => |r| try self.addRect(.xz_rect, r.x0, r.x1, r.z0, r.z1, r.k, r.material.get(), chain),
            .yzRect => |r| try self.addRect(.yz_rect, r.y0, r.y1, r.z0, r.z1, r.k, r.material.get(), chain),
            .box => |b| for (b.sides.objects.items) |o| try self.walk(o, chain),
            .translate => |t| {
                try self.xforms.append(.{ .kind = @intFromEnum(abi.XformKind.translate), .outer = chain, .v = .{ t.offset.x, t.offset.y, t.offset.z, 0 } });
                try self.walk(t.object.get().*, @intCast(self.xforms.items.len - 1));
            },
            .rotateY => |r| {
                try self.xforms.append(.{ .kind = @intFromEnum(abi.XformKind.rotate_y), .outer = chain, .v = .{ r.sin_t, r.cos_t, 0, 0 } });
                try self.walk(r.object.get().*, @intCast(self.xforms.items.len - 1));
            },
        }
    }
};

pub fn flatten(allocator: std.mem.Allocator, world: Hittable, time0: f64, time1: f64) !FlatScene {
    var fs = FlatScene.init(allocator, time0, time1);
    errdefer fs.deinit();
    try fs.walk(world, -1);
    return fs;
}
