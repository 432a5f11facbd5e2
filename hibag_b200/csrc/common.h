// common.h -- error handling and RAII device / pinned buffers (internal)
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/hibag_b200.h"

namespace hb {

#define HB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
		" at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

/// selected device + its properties; throws when no CUDA device is usable (there is no CPU
/// fallback in this library)
struct DeviceInfo
{
	int device;
	int sm_count;
	int clock_khz;
	char name[128];
};
const DeviceInfo &current_device();
void select_device(int device);

/// Process-wide cache of device and page-locked host blocks. cudaMalloc / cudaMallocHost /
/// cudaFree synchronise the device (and serialise across processes for pinned memory); training
/// sessions and predict calls come and go, so released blocks are kept and handed out again.
void *pool_alloc(bool pinned, size_t bytes, size_t *got_bytes);
void pool_free(bool pinned, void *p, size_t bytes);
/// give every cached block back to the driver (returns the bytes released)
size_t pool_trim();

/// growable device buffer (contents are NOT preserved on growth)
template <typename T>
class DevBuf
{
public:
	DevBuf() : p_(nullptr), cap_(0), bytes_(0) {}
	~DevBuf() { release(); }
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	T *ensure(size_t n)
	{
		if (n > cap_)
		{
			// the old block goes back to the process-wide cache, where any other thread may take it at
			// once: nothing enqueued earlier (on any stream, e.g. an async predict call) may still use it
			if (p_) cudaDeviceSynchronize();
			release();
			const size_t want = n + n / 4 + 64;
			p_ = (T *)pool_alloc(false, want * sizeof(T), &bytes_);
			cap_ = bytes_ / sizeof(T);
		}
		return p_;
	}
	T *get() const { return p_; }
	size_t capacity() const { return cap_; }
	void release() { if (p_) { pool_free(false, p_, bytes_); p_ = nullptr; cap_ = 0; bytes_ = 0; } }
private:
	T *p_;
	size_t cap_, bytes_;
};

/// growable page-locked host buffer
template <typename T>
class PinBuf
{
public:
	PinBuf() : p_(nullptr), cap_(0), bytes_(0) {}
	~PinBuf() { release(); }
	PinBuf(const PinBuf &) = delete;
	PinBuf &operator=(const PinBuf &) = delete;
	T *ensure(size_t n)
	{
		if (n > cap_)
		{
			release();
			const size_t want = n + n / 4 + 64;
			p_ = (T *)pool_alloc(true, want * sizeof(T), &bytes_);
			cap_ = bytes_ / sizeof(T);
		}
		return p_;
	}
	T *get() const { return p_; }
	void release() { if (p_) { pool_free(true, p_, bytes_); p_ = nullptr; cap_ = 0; bytes_ = 0; } }
private:
	T *p_;
	size_t cap_, bytes_;
};

struct Stream
{
	cudaStream_t s = nullptr;
	/// high_priority: the stream's kernels are dispatched before pending work of default-priority
	/// streams (short scoring launches ahead of the long, SM-filling EM launches)
	explicit Stream(bool high_priority = false)
	{
		if (high_priority)
		{
			int lo = 0, hi = 0;
			HB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
			HB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
		} else
			HB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	}
	~Stream() { if (s) cudaStreamDestroy(s); }
	Stream(const Stream &) = delete;
	Stream &operator=(const Stream &) = delete;
};

struct Event
{
	cudaEvent_t e = nullptr;
	explicit Event(bool timing = true, bool blocking = false)
	{
		unsigned flags = timing ? cudaEventDefault : cudaEventDisableTiming;
		if (blocking) flags |= cudaEventBlockingSync;    // waiting host thread sleeps instead of spinning
		HB_CUDA(cudaEventCreateWithFlags(&e, flags));
	}
	~Event() { if (e) cudaEventDestroy(e); }
	Event(const Event &) = delete;
	Event &operator=(const Event &) = delete;
};

/// Wait for a stream WITHOUT spinning: cudaStreamSynchronize busy-waits under the default
/// scheduling policy, and a trainer runs dozens of lane threads on a handful of cores (measured:
/// 6.4 cores busy at 40 lanes, most of it in the two synchronisations of RoundEM::prepare).
inline void stream_sync_blocking(cudaStream_t s)
{
	thread_local Event ev(false, true);
	HB_CUDA(cudaEventRecord(ev.e, s));
	HB_CUDA(cudaEventSynchronize(ev.e));
}

}  // namespace hb
