// NOT COMPILED OR RUN: this image has no cargo/rustc.  Written against the reference's sources (bvpav/raydar) and
// include/raydar_cuda.h; the same C ABI is exercised through ctypes by tests/ and bench.py.  See INTEGRATION.md.
//! `impl Renderer for CudaRenderer` -- drop into the reference as src/renderer/cuda.rs (renderer/mod.rs:25-35)
use image::RgbaImage;
use raydar_cuda_sys as sys;
use super::{timing::Profiler, Renderer, RendererConfig};
use crate::scene::{objects::Geometry, world::World, Scene};

pub struct CudaRenderer {
    handle: *mut sys::RdrRenderer,
    profiler: Profiler,
    resolution: (u32, u32),
    /// pinned image the GPUs write directly (rdr_alloc_host_image), re-allocated when the resolution changes
    image: *mut u8,
    image_bytes: usize,
}

/// Flattened, borrowed view of `&Scene` (the Vec's must outlive the FFI call).
struct Flat { kind: Vec<u32>, geom: Vec<f32>, material: Vec<f32>, raw: sys::RdrSceneFlat }

fn flatten(scene: &Scene) -> Flat {
    let cam = &scene.camera;
    let mut kind = Vec::new(); let mut geom = Vec::new(); let mut material = Vec::new();
    for o in &scene.objects {
        match &o.geometry {
            Geometry::Sphere(s) => { kind.push(0); geom.extend([s.center.x, s.center.y, s.center.z, s.radius]); }
            Geometry::Cube(c)   => { kind.push(1); geom.extend([c.center.x, c.center.y, c.center.z, c.side_length]); }
        }
        let m = &o.material;
        material.extend([m.albedo.x, m.albedo.y, m.albedo.z, m.roughness, m.metallic,
                         m.emission_color.x, m.emission_color.y, m.emission_color.z,
                         m.emission_strength, m.transmission, m.ior]);
    }
    let (world_kind, a, b) = match &scene.world {
        World::SkyColor { top_color, bottom_color } => (0, *top_color, *bottom_color),
        World::SolidColor(c) => (1, *c, cgmath::Vector3::new(0.0, 0.0, 0.0)),
        World::Transparent => (2, cgmath::Vector3::new(0.0, 0.0, 0.0), cgmath::Vector3::new(0.0, 0.0, 0.0)),
    };
    let m4 = |m: cgmath::Matrix4<f32>| -> [f32; 16] { *AsRef::<[f32; 16]>::as_ref(&m) };    // column-major
    let p = cam.position();
    let raw = sys::RdrSceneFlat {
        width: cam.resolution_x(), height: cam.resolution_y(),
        inv_proj: m4(cam.inverse_proj_matrix()), inv_view: m4(cam.inverse_view_matrix()),
        cam_pos: [p.x, p.y, p.z], world_kind, world_a: [a.x, a.y, a.z], world_b: [b.x, b.y, b.z],
        n_objects: kind.len() as u32, kind: kind.as_ptr(), geom: geom.as_ptr(), material: material.as_ptr(),
    };
    Flat { kind, geom, material, raw }
}

impl CudaRenderer {
    pub fn new(config: RendererConfig) -> Self {                        // by convention, cpu.rs:186 / vulkan.rs:458
        let cfg = sys::RdrConfig { max_sample_count: config.max_sample_count, max_bounces: config.max_bounces };
        let mut handle = std::ptr::null_mut();
        let st = unsafe { sys::rdr_create(&cfg, 0, &mut handle) };
        assert!(st == 0, "rdr_create failed: {}", last_error(std::ptr::null()));
        Self { handle, profiler: Profiler::default(), resolution: (0, 0), image: std::ptr::null_mut(), image_bytes: 0 }
    }
    /// `--gpus N`: one sub-renderer per device behind the same handle (rdr_create_multi).  `stripes`: round-robin
    /// 16-row stripes instead of sample ranges -- the image is then bit-identical to the one-GPU image.
    pub fn with_gpus(config: RendererConfig, gpus: u32, stripes: bool) -> Self {
        let cfg = sys::RdrConfig { max_sample_count: config.max_sample_count, max_bounces: config.max_bounces };
        let devices: Vec<std::os::raw::c_int> = (0..gpus as std::os::raw::c_int).collect();
        let mut handle = std::ptr::null_mut();
        let st = unsafe { sys::rdr_create_multi(&cfg, devices.len() as std::os::raw::c_int, devices.as_ptr(), &mut handle) };
        assert!(st == 0, "rdr_create_multi failed: {}", last_error(std::ptr::null()));
        if stripes { assert!(unsafe { sys::rdr_set_partition(handle, 1, 16) } == 0, "rdr_set_partition failed"); }
        Self { handle, profiler: Profiler::default(), resolution: (0, 0), image: std::ptr::null_mut(), image_bytes: 0 }
    }
    fn check(&self, st: i32) { assert!(st == 0, "raydar_cuda: {}", last_error(self.handle)); }   // the trait is infallible
    fn image_for(&mut self, w: u32, h: u32) -> *mut u8 {
        let bytes = (w as usize) * (h as usize) * 4;
        if bytes > self.image_bytes {
            unsafe { sys::rdr_free_host_image(self.image) };
            self.image = std::ptr::null_mut(); self.image_bytes = 0;
            assert!(unsafe { sys::rdr_alloc_host_image(bytes.max(4), &mut self.image) } == 0, "rdr_alloc_host_image failed");
            self.image_bytes = bytes.max(4);
        }
        self.image
    }
    fn take_image(&self, w: u32, h: u32) -> RgbaImage {        // the trait returns an owned image: one host copy of W*H*4 bytes
        let n = (w as usize) * (h as usize) * 4;
        RgbaImage::from_raw(w, h, unsafe { std::slice::from_raw_parts(self.image, n) }.to_vec()).unwrap()
    }
}

// The Profiler's timers are `pub(super)` (timing.rs:12-18), so this sibling module drives them exactly as cpu.rs and
// vulkan.rs do -- through Timer's public start() / end() / end_multiple() around the blocking FFI calls -- and needs no
// change to timing.rs.  (The library keeps its own copy of the four durations plus the CUDA-event kernel time:
// rdr_profiler, used by the C++ CLI and bench.py.)
impl Renderer for CudaRenderer {
    fn render_frame(&mut self, scene: &Scene) -> RgbaImage {
        self.new_frame(scene);                                               // cpu.rs:120: frame + prepare timers start
        let (w, h) = self.resolution;
        let img = self.image_for(w, h);
        let n = self.max_sample_count();
        self.profiler.prepare_timer.end_if_not_ended();                      // cpu.rs:194
        self.profiler.render_timer.start();
        self.profiler.sample_timer.start();
        let st = unsafe { sys::rdr_finish_frame(self.handle, img) };         // all samples + resolve, blocking (vulkan.rs:168)
        self.check(st);
        if n > 0 {                                                           // n == 0: the timers stay unset, as in cpu.rs
            self.profiler.sample_timer.end_multiple(n);                      // vulkan.rs:175-179: mean per-sample time
            self.profiler.render_timer.end();
            self.profiler.frame_timer.end();
        }
        self.take_image(w, h)
    }
    fn new_frame(&mut self, scene: &Scene) {
        self.profiler.frame_timer.start();                                   // cpu.rs:136-137
        self.profiler.prepare_timer.start();
        let flat = flatten(scene);
        let st = unsafe { sys::rdr_new_frame(self.handle, &flat.raw) };      // snapshots the scene (vulkan.rs:207-428)
        self.check(st); self.resolution = (flat.raw.width, flat.raw.height);
        self.profiler.prepare_timer.end();
    }
    fn render_sample(&mut self, scene: &Scene) -> Option<RgbaImage> {
        if self.resolution == (0, 0) { self.new_frame(scene); }              // the reference allocates its frame buffer lazily (cpu.rs:400-411)
        let (w, h) = self.resolution;
        let img = self.image_for(w, h);
        let mut produced = 0;
        self.profiler.render_timer.start_if_not_started();                   // cpu.rs:195
        self.profiler.sample_timer.start();                                  // cpu.rs:196
        let st = unsafe { sys::rdr_render_sample(self.handle, img, &mut produced) };
        self.check(st);
        if produced == 0 { return None; }                                    // cpu.rs:143-145
        self.profiler.sample_timer.end();                                    // cpu.rs:218
        if self.sample_count() == self.max_sample_count() {                  // cpu.rs:213-216
            self.profiler.render_timer.end();
            self.profiler.frame_timer.end();
        }
        Some(self.take_image(w, h))
    }
    fn profiler(&self) -> &Profiler { &self.profiler }
    fn sample_count(&self) -> u32 { unsafe { sys::rdr_sample_count(self.handle) } }
    fn max_sample_count(&self) -> u32 { unsafe { sys::rdr_max_sample_count(self.handle) } }
    fn max_bounces(&self) -> u32 { unsafe { sys::rdr_max_bounces(self.handle) } }
    fn set_max_sample_count(&mut self, count: u32) { unsafe { sys::rdr_set_max_sample_count(self.handle, count) }; }
    fn set_max_bounces(&mut self, bounces: u32) { unsafe { sys::rdr_set_max_bounces(self.handle, bounces) }; }
}

impl Drop for CudaRenderer {
    fn drop(&mut self) { unsafe { sys::rdr_free_host_image(self.image); sys::rdr_destroy(self.handle) } }
}

fn last_error(h: *const sys::RdrRenderer) -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::rdr_last_error(h)).to_string_lossy().into_owned() }
}
