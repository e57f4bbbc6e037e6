// rdr_loop.cuh -- shared-memory columns of the cold / parked lane state of the render kernel's sample loop (the loop
// body itself is rdr_loop_body.inc).  Also compiled for the CPU by tests/hostsim (RDR_WARP_EMU).
#pragma once

#include "rdr_trace.cuh"

namespace rdr {

// Cold per-lane state (accumulator, primary direction, primary hit: LaneStateT in rdr_trace.cuh) in shared memory:
// word i of the lane at base[i * BLOCK + threadIdx.x] (conflict-free columns).  The COLD variants of the warp-cooperative
// kernels run their hot loop with 9 registers fewer, and with 18 fewer inside the nearest-hit search (lane_park), which is
// what lets a 896-thread CTA (28 warps per SM at 72 registers) run without spills.
template <int BLOCK>
struct ColdShared {
    float *col;             // &base[threadIdx.x]
    __device__ __forceinline__ float get(int i) const { return col[i * BLOCK]; }
    __device__ __forceinline__ void set(int i, float v) { col[i * BLOCK] = v; }
};
template <int BLOCK, bool COLD> struct lane_state_of { typedef LaneState type; };
template <int BLOCK> struct lane_state_of<BLOCK, true> { typedef LaneStateT<ColdShared<BLOCK> > type; };
// the path state the scan does not need (throughput, light, RNG counters: 9 words) waits in the lane's shared-memory
// columns while the warp is inside the nearest-hit search
template <class ST> __device__ __forceinline__ void lane_park(ST &) {}
template <class ST> __device__ __forceinline__ void lane_unpark(ST &) {}
template <int BLOCK> __device__ __forceinline__ void lane_park(LaneStateT<ColdShared<BLOCK> > &st)
{
    st.cold.set(COLD_PARK + 0, st.light.x); st.cold.set(COLD_PARK + 1, st.light.y); st.cold.set(COLD_PARK + 2, st.light.z);
    st.cold.set(COLD_PARK + 3, st.atten.x); st.cold.set(COLD_PARK + 4, st.atten.y); st.cold.set(COLD_PARK + 5, st.atten.z);
    st.cold.set(COLD_PARK + 6, u2f(st.pixel)); st.cold.set(COLD_PARK + 7, u2f(st.s)); st.cold.set(COLD_PARK + 8, u2f(st.bounce));
}
template <int BLOCK> __device__ __forceinline__ void lane_unpark(LaneStateT<ColdShared<BLOCK> > &st)
{
    st.light = mk3(st.cold.get(COLD_PARK + 0), st.cold.get(COLD_PARK + 1), st.cold.get(COLD_PARK + 2));
    st.atten = mk3(st.cold.get(COLD_PARK + 3), st.cold.get(COLD_PARK + 4), st.cold.get(COLD_PARK + 5));
    st.pixel = f2u(st.cold.get(COLD_PARK + 6)); st.s = f2u(st.cold.get(COLD_PARK + 7)); st.bounce = f2u(st.cold.get(COLD_PARK + 8));
}
__device__ __forceinline__ void cold_bind(ColdRegs &, unsigned char *) {}
template <int BLOCK> __device__ __forceinline__ void cold_bind(ColdShared<BLOCK> &c, unsigned char *base) { c.col = reinterpret_cast<float *>(base) + threadIdx.x; }

}  // namespace rdr
