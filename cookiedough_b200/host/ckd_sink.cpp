// ckd_sink.cpp -- headless frame sink: what takes the finished frame where the reference hands it to
// Display::Update (display.cpp:66-82, called from main.cpp:336-345).  Frames are written to one raw stream file
//
//     "CKDF" u32 version(1) u32 resX u32 resY u32 numFrames u32 reserved[3]      (32-byte header)
//     numFrames x resX*resY little-endian 0xAARRGGBB pixels                      (frame i at 32 + i*resX*resY*4)
//
// by a writer thread, through a ring of (pinned) host buffers: the renderer acquires a buffer, lets X_Draw / Demo_Draw fill
// it, commits it with its frame index and carries on with the next frame while the write is in flight.  Frames are placed by
// index (pwrite), so they may be committed in any order and by several processes (one per GPU) into the same file.

#include "../../include/ckd_host.h"

#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Sink
{
	int fd = -1;
	unsigned resX = 0, resY = 0, numFrames = 0;
	size_t frameBytes = 0;
	bool pinned = false;
	std::vector<uint32_t *> buffers;

	std::mutex mutex;
	std::condition_variable freed, queued;
	std::deque<uint32_t *> freeList;
	std::deque<std::pair<uint32_t *, unsigned>> pending;
	bool closing = false;
	std::string error;
	unsigned long long written = 0;
	std::thread writer;
};

Sink *s_sink = nullptr;
constexpr size_t kHeaderBytes = 32;

void WriterLoop(Sink *sink)
{
	for (;;)
	{
		std::pair<uint32_t *, unsigned> job;
		{
			std::unique_lock<std::mutex> lock(sink->mutex);
			sink->queued.wait(lock, [&] { return !sink->pending.empty() || sink->closing; });
			if (sink->pending.empty())
				return;
			job = sink->pending.front();
			sink->pending.pop_front();
		}

		const uint8_t *p = reinterpret_cast<const uint8_t *>(job.first);
		size_t left = sink->frameBytes;
		off_t offset = off_t(kHeaderBytes) + off_t(job.second)*off_t(sink->frameBytes);
		std::string error;
		while (left > 0)
		{
			const ssize_t n = pwrite(sink->fd, p, left, offset);
			if (n < 0)
			{
				if (EINTR == errno) continue;
				error = std::string("CkdSink: write failed: ") + strerror(errno);
				break;
			}
			p += n; left -= size_t(n); offset += n;
		}

		std::lock_guard<std::mutex> lock(sink->mutex);
		if (!error.empty() && sink->error.empty()) sink->error = error;
		++sink->written;
		sink->freeList.push_back(job.first);
		sink->freed.notify_all();
	}
}

} // namespace

// opens (creating it when 'create' is set: one process does that, the others attach) the stream for numFrames frames
bool CkdSink_Open(const char *path, unsigned resX, unsigned resY, unsigned numFrames, unsigned ringFrames, bool pinned, bool create)
{
	if (nullptr != s_sink)
	{
		SetLastError("CkdSink_Open: a sink is already open");
		return false;
	}
	if (nullptr == path || 0 == resX || 0 == resY || 0 == ringFrames)
	{
		SetLastError("CkdSink_Open: invalid argument");
		return false;
	}

	Sink *sink = new Sink;
	sink->resX = resX; sink->resY = resY; sink->numFrames = numFrames;
	sink->frameBytes = size_t(resX)*resY*4;
	sink->pinned = pinned;
	sink->fd = open(path, create ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
	if (sink->fd < 0)
	{
		SetLastError(std::string("CkdSink_Open: ") + path + ": " + strerror(errno));
		delete sink;
		return false;
	}
	if (create)
	{
		uint32_t header[8] = { 0x46444b43u /* "CKDF" */, 1u, resX, resY, numFrames, 0u, 0u, 0u };
		if (pwrite(sink->fd, header, sizeof(header), 0) != ssize_t(sizeof(header))
			|| 0 != ftruncate(sink->fd, off_t(kHeaderBytes) + off_t(numFrames)*off_t(sink->frameBytes)))
		{
			SetLastError(std::string("CkdSink_Open: ") + path + ": " + strerror(errno));
			close(sink->fd);
			delete sink;
			return false;
		}
	}

	else
	{
		// attaching: the stream must be the one this process is about to write into
		uint32_t header[8] = { 0 };
		const int rfd = open(path, O_RDONLY);
		const bool readOk = rfd >= 0 && pread(rfd, header, sizeof(header), 0) == ssize_t(sizeof(header));
		if (rfd >= 0) close(rfd);
		if (!readOk || 0x46444b43u != header[0] || 1u != header[1] || resX != header[2] || resY != header[3] || numFrames != header[4])
		{
			SetLastError(std::string("CkdSink_Open: ") + path + " is not a CKDF stream of this resolution and frame count");
			close(sink->fd);
			delete sink;
			return false;
		}
	}

	for (unsigned i = 0; i < ringFrames; ++i)
	{
		void *p = nullptr;
		if (pinned)
		{
			if (CKD_OK != ckd_malloc_host(&p, sink->frameBytes)) p = nullptr;
		}
		else if (0 != posix_memalign(&p, 4096, sink->frameBytes))
			p = nullptr;
		if (nullptr == p)
		{
			SetLastError("CkdSink_Open: out of host memory for the frame ring");
			for (uint32_t *b : sink->buffers) { if (pinned) ckd_free_host(b); else free(b); }
			close(sink->fd);
			delete sink;
			return false;
		}
		sink->buffers.push_back(static_cast<uint32_t *>(p));
		sink->freeList.push_back(static_cast<uint32_t *>(p));
	}

	sink->writer = std::thread(WriterLoop, sink);
	s_sink = sink;
	return true;
}

// next free frame buffer of the ring; blocks while every buffer is still being written
uint32_t *CkdSink_Acquire()
{
	Sink *sink = s_sink;
	if (nullptr == sink)
	{
		SetLastError("CkdSink_Acquire: no sink is open");
		return nullptr;
	}
	std::unique_lock<std::mutex> lock(sink->mutex);
	sink->freed.wait(lock, [&] { return !sink->freeList.empty(); });
	uint32_t *p = sink->freeList.front();
	sink->freeList.pop_front();
	return p;
}

// hands a filled buffer (from CkdSink_Acquire) to the writer as frame 'frameIndex'
bool CkdSink_Commit(uint32_t *frame, unsigned frameIndex)
{
	Sink *sink = s_sink;
	if (nullptr == sink || nullptr == frame)
	{
		SetLastError("CkdSink_Commit: invalid argument");
		return false;
	}
	std::lock_guard<std::mutex> lock(sink->mutex);
	if (frameIndex >= sink->numFrames || !sink->error.empty())
	{
		// a rejected frame's buffer goes back to the ring (otherwise ringFrames rejections would block CkdSink_Acquire for good)
		if (frameIndex >= sink->numFrames) SetLastError("CkdSink_Commit: frame index past the end of the stream");
		else SetLastError(sink->error);
		sink->freeList.push_back(frame);
		sink->freed.notify_all();
		return false;
	}
	sink->pending.emplace_back(frame, frameIndex);
	sink->queued.notify_one();
	return true;
}

// waits for all committed frames, closes the file; false when any write failed
bool CkdSink_Close()
{
	Sink *sink = s_sink;
	if (nullptr == sink)
		return true;
	{
		std::lock_guard<std::mutex> lock(sink->mutex);
		sink->closing = true;
		sink->queued.notify_all();
	}
	sink->writer.join();
	const bool ok = sink->error.empty() && 0 == fsync(sink->fd);
	if (!sink->error.empty()) SetLastError(sink->error);
	close(sink->fd);
	for (uint32_t *b : sink->buffers) { if (sink->pinned) ckd_free_host(b); else free(b); }
	delete sink;
	s_sink = nullptr;
	return ok;
}

extern "C" {

int ckdsink_open(const char *path, unsigned resX, unsigned resY, unsigned numFrames, unsigned ringFrames, int pinned, int create)
{
	return CkdSink_Open(path, resX, resY, numFrames, ringFrames, 0 != pinned, 0 != create) ? 0 : -1;
}
uint32_t *ckdsink_acquire() { return CkdSink_Acquire(); }
int ckdsink_commit(uint32_t *frame, unsigned frameIndex) { return CkdSink_Commit(frame, frameIndex) ? 0 : -1; }
int ckdsink_close() { return CkdSink_Close() ? 0 : -1; }

} // extern "C"
