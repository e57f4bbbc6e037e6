// NOT COMPILED OR RUN: this image has no cargo/rustc.  Written against the reference's sources (bvpav/raydar) and
// include/raydar_cuda.h; the same C ABI is exercised through ctypes by tests/ and bench.py.  See INTEGRATION.md.
//! raw bindings of libraydar_cuda.so (include/raydar_cuda.h)
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct RdrRenderer { _private: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy)]
pub struct RdrConfig { pub max_sample_count: u32, pub max_bounces: u32 }          // renderer/mod.rs:11-14

#[repr(C)]
pub struct RdrSceneFlat {
    pub width: u32, pub height: u32,
    pub inv_proj: [f32; 16], pub inv_view: [f32; 16],       // column-major, as cgmath stores Matrix4
    pub cam_pos: [f32; 3],
    pub world_kind: u32,                                     // 0 SkyColor, 1 SolidColor, 2 Transparent
    pub world_a: [f32; 3], pub world_b: [f32; 3],
    pub n_objects: u32,
    pub kind: *const u32,                                    // 0 Sphere, 1 Cube
    pub geom: *const f32,                                    // n * (center.xyz, radius | side_length)
    pub material: *const f32,                                // n * 11, field order of scene/material.rs:4-13
}

#[repr(C)] #[derive(Default)]
pub struct RdrProfiler {
    pub frame_ns: u64, pub sample_ns: u64, pub prepare_ns: u64, pub render_ns: u64,
    pub has_frame: u32, pub has_sample: u32, pub has_prepare: u32, pub has_render: u32,
    pub device_render_ms: f64,
}

#[link(name = "raydar_cuda")]
extern "C" {
    pub fn rdr_create(config: *const RdrConfig, device: c_int, out: *mut *mut RdrRenderer) -> c_int;
    pub fn rdr_create_multi(config: *const RdrConfig, n: c_int, devices: *const c_int, out: *mut *mut RdrRenderer) -> c_int;
    pub fn rdr_destroy(r: *mut RdrRenderer);
    pub fn rdr_last_error(r: *const RdrRenderer) -> *const c_char;
    pub fn rdr_new_frame(r: *mut RdrRenderer, scene: *const RdrSceneFlat) -> c_int;
    pub fn rdr_render_sample(r: *mut RdrRenderer, rgba8: *mut u8, produced: *mut c_int) -> c_int;
    pub fn rdr_render_frame(r: *mut RdrRenderer, scene: *const RdrSceneFlat, rgba8: *mut u8) -> c_int;
    /// the second half of render_frame: every sample the current frame has left + the image
    pub fn rdr_finish_frame(r: *mut RdrRenderer, rgba8: *mut u8) -> c_int;
    pub fn rdr_profiler(r: *const RdrRenderer, out: *mut RdrProfiler) -> c_int;
    pub fn rdr_sample_count(r: *const RdrRenderer) -> u32;
    pub fn rdr_max_sample_count(r: *const RdrRenderer) -> u32;
    pub fn rdr_max_bounces(r: *const RdrRenderer) -> u32;
    pub fn rdr_set_max_sample_count(r: *mut RdrRenderer, count: u32) -> c_int;
    pub fn rdr_set_max_bounces(r: *mut RdrRenderer, bounces: u32) -> c_int;
    pub fn rdr_set_seed(r: *mut RdrRenderer, seed: u64) -> c_int;
    /// multi-GPU handle: 0 = sample ranges (default), 1 = round-robin row stripes (bit-identical to one GPU)
    pub fn rdr_set_partition(r: *mut RdrRenderer, partition: c_int, stripe_rows: u32) -> c_int;
    /// pinned host image buffers: the GPUs write an image that lies in one directly (no staging copy)
    pub fn rdr_alloc_host_image(bytes: usize, out: *mut *mut u8) -> c_int;
    pub fn rdr_free_host_image(image: *mut u8);
}
