// NOT COMPILED OR RUN: this image has no cargo/rustc.  Written against the reference's sources (bvpav/raydar) and
// include/raydar_cuda.h; the same C ABI is exercised through ctypes by tests/ and bench.py.  See INTEGRATION.md.
//! `impl Renderer for CudaRenderer` -- drop into the reference as src/renderer/cuda.rs (renderer/mod.rs:25-35)
use image::RgbaImage;
use raydar_cuda_sys as sys;
use super::{timing::Profiler, Renderer, RendererConfig};
use crate::scene::{objects::Geometry, world::World, Scene};

pub struct CudaRenderer { handle: *mut sys::RdrRenderer, profiler: Profiler, resolution: (u32, u32) }

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
        Self { handle, profiler: Profiler::default(), resolution: (0, 0) }
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
        Self { handle, profiler: Profiler::default(), resolution: (0, 0) }
    }
    fn check(&self, st: i32) { assert!(st == 0, "raydar_cuda: {}", last_error(self.handle)); }   // the trait is infallible
    fn sync_profiler(&mut self) {        // timers are pub(super): a sibling module may fill them (timing.rs:12-18)
        let mut p = sys::RdrProfiler::default();
        unsafe { sys::rdr_profiler(self.handle, &mut p) };
        self.profiler.set_durations_ns(p.has_frame != 0, p.frame_ns, p.has_sample != 0, p.sample_ns,
                                       p.has_prepare != 0, p.prepare_ns, p.has_render != 0, p.render_ns);
        // `set_durations_ns` is a 6-line helper to add to timing.rs (Timer { duration: Some(Duration::from_nanos(..)) })
    }
}

impl Renderer for CudaRenderer {
    fn render_frame(&mut self, scene: &Scene) -> RgbaImage {
        let flat = flatten(scene);
        let (w, h) = (flat.raw.width, flat.raw.height);
        let mut buf = vec![0u8; (w * h * 4) as usize];
        let st = unsafe { sys::rdr_render_frame(self.handle, &flat.raw, buf.as_mut_ptr()) };
        self.check(st); self.sync_profiler(); self.resolution = (w, h);
        RgbaImage::from_raw(w, h, buf).unwrap()
    }
    fn new_frame(&mut self, scene: &Scene) {
        let flat = flatten(scene);
        let st = unsafe { sys::rdr_new_frame(self.handle, &flat.raw) };
        self.check(st); self.resolution = (flat.raw.width, flat.raw.height);
    }
    fn render_sample(&mut self, _scene: &Scene) -> Option<RgbaImage> {
        let (w, h) = self.resolution;
        let mut buf = vec![0u8; (w * h * 4) as usize];
        let mut produced = 0;
        let st = unsafe { sys::rdr_render_sample(self.handle, buf.as_mut_ptr(), &mut produced) };
        self.check(st); self.sync_profiler();
        if produced != 0 { RgbaImage::from_raw(w, h, buf) } else { None }
    }
    fn profiler(&self) -> &Profiler { &self.profiler }
    fn sample_count(&self) -> u32 { unsafe { sys::rdr_sample_count(self.handle) } }
    fn max_sample_count(&self) -> u32 { unsafe { sys::rdr_max_sample_count(self.handle) } }
    fn max_bounces(&self) -> u32 { unsafe { sys::rdr_max_bounces(self.handle) } }
    fn set_max_sample_count(&mut self, count: u32) { unsafe { sys::rdr_set_max_sample_count(self.handle, count) }; }
    fn set_max_bounces(&mut self, bounces: u32) { unsafe { sys::rdr_set_max_bounces(self.handle, bounces) }; }
}

impl Drop for CudaRenderer { fn drop(&mut self) { unsafe { sys::rdr_destroy(self.handle) } } }

fn last_error(h: *const sys::RdrRenderer) -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::rdr_last_error(h)).to_string_lossy().into_owned() }
}
