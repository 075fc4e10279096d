// starway_b200 — layouts shared by the host progress engine and the sm_100a kernels.
//
// Everything in this file is plain-old-data that lives either in device memory
// (rings, queues, job lists), in pinned host memory (per-launch input/output
// blocks) or in POSIX shared memory (control block between two workers).
//
// Reference being replaced: the wire/queue state that libucp keeps behind
// ucp_tag_send_nbx / ucp_tag_recv_nbx / ucp_worker_progress
// (reference src/bindings/main.cpp:370,404,362 and :1136,1172,1127).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SW_HD __host__ __device__
#else
#define SW_HD
#endif

// ---------------------------------------------------------------- inbound ring
// One ring per inbound endpoint, owned by the receiving worker, written by the
// sending worker's put kernel through a peer mapping (NVLink stores) or a plain
// pointer (same process).  A slot is a fixed 8 KiB landing zone.
constexpr uint32_t SW_SLOT_BYTES = 8192;
constexpr uint32_t SW_SLOT_HDR = 64;
constexpr uint32_t SW_EAGER_MAX = SW_SLOT_BYTES - SW_SLOT_HDR;  // 8128 B payload
constexpr uint32_t SW_RING_SLOTS_DEFAULT = 1024;                // 8 MiB per endpoint

enum : uint32_t { SW_KIND_EAGER = 1, SW_KIND_RTS = 2 };

struct SwSlotHdr {     // 64 B, 16 B aligned, written last by the put kernel
  uint64_t tag;        // sender tag (full uint64)
  uint64_t len;        // message length in bytes (for RTS: length of the remote payload)
  uint64_t seq;        // 1-based slot sequence number on this ring (sanity flag)
  uint32_t kind;       // SW_KIND_*
  uint32_t magic;      // SW_SLOT_MAGIC
  uint64_t pad[4];
};
constexpr uint32_t SW_SLOT_MAGIC = 0x53574d47u;  // "SWMG"

// Rendezvous request-to-send descriptor; travels as the payload of an RTS slot so
// that it stays ordered with the eager messages of the same sender (non-overtaking).
struct SwRts {            // 128 B
  uint8_t ipc_handle[64];  // cudaIpcMemHandle_t of the allocation holding the source
  uint64_t alloc_base;     // sender-side virtual address of the allocation base
  uint64_t alloc_size;     // size of that allocation
  uint64_t src_ptr;        // sender-side virtual address of the payload
  uint64_t send_seq;       // sender's rendezvous sequence number (echoed in FIN)
  uint64_t ctx_uuid;       // sender context id (same value => same process => direct ptr)
  uint32_t src_pid;
  int32_t src_dev;         // CUDA ordinal of the source memory (-1: unknown)
  uint64_t pad[2];
};
static_assert(sizeof(SwRts) == 128, "SwRts layout");
static_assert(sizeof(SwSlotHdr) == 64, "SwSlotHdr layout");

// ---------------------------------------------------------------- put kernel input
struct SwPutDesc {   // pinned host, read by the put kernel
  uint64_t src;      // payload source (device ptr, or pinned-host staging ptr)
  uint64_t dst;      // destination slot base (peer-mapped or local device ptr)
  uint64_t tag;
  uint64_t seq;      // ring sequence number for the slot header
  uint32_t len;      // payload bytes to copy into the slot (<= SW_EAGER_MAX)
  uint32_t kind;     // SW_KIND_*
  uint64_t msg_len;  // value for SwSlotHdr.len (== len for eager; remote length for RTS)
};
static_assert(sizeof(SwPutDesc) == 48, "SwPutDesc layout");

// ---------------------------------------------------------------- match kernel I/O
constexpr uint32_t SW_MAX_EPS = 64;          // inbound endpoints per worker
constexpr uint32_t SW_MAX_POSTS = 1024;      // new receives per match launch
constexpr uint32_t SW_MAX_ARRIVALS = 1024;   // ring slots consumed per match launch
constexpr uint32_t SW_MAX_JOBS = SW_MAX_POSTS + SW_MAX_ARRIVALS;

// Completion status codes (values mirror the UCX codes the reference forwards
// through ucs_status_string; see sw_status_string in the C-ABI).
enum : int32_t {
  SW_OK = 0,
  SW_ERR_MESSAGE_TRUNCATED = -9,
  SW_ERR_CANCELED = -16,
  SW_ERR_NOT_CONNECTED = -24,
  SW_ERR_CONNECTION_RESET = -25,
  SW_ERR_INVALID_PARAM = -5,
  SW_ERR_NO_MEMORY = -4,
  SW_ERR_IO_ERROR = -3,
  SW_ERR_UNREACHABLE = -6,
  SW_ERR_TIMED_OUT = -20,
  SW_ERR_BUSY = -15,
};

struct SwPost {     // one posted receive (reference ClientRecvArgs/ServerRecvArgs, main.hpp:80-86,234-240)
  uint64_t tag;
  uint64_t mask;
  uint64_t buf;     // device-visible destination
  uint64_t cap;
  uint64_t op_id;
};

struct SwMatchIn {  // pinned host
  uint32_t n_posts;
  uint32_t n_eps;
  uint32_t max_arrivals;   // budget for this launch (<= SW_MAX_ARRIVALS)
  uint32_t pad;
  uint64_t produced[SW_MAX_EPS];  // slots produced so far on each inbound ring
  SwPost posts[SW_MAX_POSTS];
};

enum : uint32_t {
  SW_JOB_DELIVER = 1,   // copy into a posted receive buffer and complete it
  SW_JOB_STASH = 2,     // copy an unexpected eager payload from its slot to the heap
};

struct SwJob {      // device memory: match kernel -> deliver kernel
  uint64_t src;
  uint64_t dst;
  uint64_t len;     // bytes to copy (0 for truncated / zero-length)
  uint64_t op_id;   // receive being completed (0 for STASH)
  uint64_t tag;     // sender tag
  uint64_t msg_len; // full message length (reported as `length`)
  int32_t status;
  uint32_t kind;
};

struct SwCqe {      // pinned host: deliver kernel -> host
  uint64_t op_id;
  uint64_t tag;
  uint64_t len;
  int32_t status;
  uint32_t kind;
};

struct SwRndvRec {  // pinned host: match kernel -> host (receiver pulls the payload)
  uint64_t op_id;
  uint64_t dst;
  uint64_t cap;
  uint64_t tag;
  uint64_t len;
  uint32_t ep;
  int32_t status;   // SW_OK or SW_ERR_MESSAGE_TRUNCATED
  SwRts rts;
};

// Queue cursors of one worker.  The match kernel is the only writer; it reports them after every
// launch (SwMatchOut::sc) and the host hands them back as kernel parameters for the next launch, so a
// launch does not start with a chain of dependent global loads.
constexpr uint32_t SW_SC_EPS = 8;
struct SwMatchScalars {
  uint64_t p_head, p_tail, u_head, u_tail;
  uint32_t p_count, u_count;
  uint32_t n_free_small, n_free_big, n_pend_small, n_pend_big;
  uint32_t rr_ep;
  uint32_t valid;                     // 0: the kernel reads the cursors from SwMatchState
  uint64_t ring_cons[SW_SC_EPS];
  uint64_t ring_base[SW_SC_EPS];
  uint32_t ring_slots[SW_SC_EPS];
};

struct SwMatchOut {  // pinned host
  uint32_t n_jobs;
  uint32_t n_rndv;
  uint32_t n_posted;      // receives left in the posted queue after this launch
  uint32_t n_unexp;       // messages left in the unexpected queue
  uint32_t err;           // device-side consistency errors (bad magic/seq)
  uint32_t n_arrivals;    // slots consumed by this launch
  uint32_t heap_small_free;
  uint32_t heap_big_free;
  uint64_t consumed[SW_MAX_EPS];  // slots consumed so far on each inbound ring
  SwMatchScalars sc;              // queue cursors after this launch
  uint64_t done_seq;              // written last by a fused match+deliver launch (system-scope release):
                                  // the host spins on it instead of on a CUDA event
  uint64_t pad_done[7];
  SwCqe cq[SW_MAX_JOBS];
  SwRndvRec rndv[SW_MAX_JOBS];
};

// ---------------------------------------------------------------- device-resident queues
// Posted-receive queue and unexpected-message queue of one worker.  Both are
// power-of-two rings addressed by monotonically increasing logical indices;
// removal from the middle leaves a tombstone (valid = 0).
constexpr uint32_t SW_PQ_CAP = 1u << 16;
constexpr uint32_t SW_UQ_CAP = 1u << 16;
constexpr uint32_t SW_HEAP_SMALL_BYTES = 256;
constexpr uint32_t SW_HEAP_BIG_BYTES = SW_EAGER_MAX;

struct SwMatchState {   // device memory, one per worker
  // posted queue (SoA)
  uint64_t p_head, p_tail;
  uint64_t* p_tag;
  uint64_t* p_mask;
  uint64_t* p_buf;
  uint64_t* p_cap;
  uint64_t* p_op;
  uint32_t* p_valid;
  uint32_t p_count;      // valid entries
  // unexpected queue (SoA)
  uint32_t u_count;
  uint64_t u_head, u_tail;
  uint64_t* u_tag;
  uint64_t* u_len;       // full message length
  uint64_t* u_data;      // heap block holding the eager payload / the SwRts
  uint32_t* u_meta;      // bit31 valid | bit30 big-block | bit29 RTS | low bits: ep
  uint32_t* u_blk;       // heap block index
  // inbound rings
  uint64_t ring_base[SW_MAX_EPS];
  uint32_t ring_slots[SW_MAX_EPS];   // power of two
  uint64_t ring_cons[SW_MAX_EPS];    // slots consumed so far
  // unexpected heap: two size classes, index stacks; frees are deferred one launch
  uint8_t* heap_small;  uint32_t* free_small;  uint32_t n_free_small;  uint32_t cap_small;
  uint8_t* heap_big;    uint32_t* free_big;    uint32_t n_free_big;    uint32_t cap_big;
  uint32_t* pend_small; uint32_t n_pend_small;
  uint32_t* pend_big;   uint32_t n_pend_big;
  // job list for the deliver kernel
  SwJob* jobs;
  uint32_t n_jobs;
  uint32_t rr_ep;        // round-robin start endpoint
};

constexpr uint32_t SW_UMETA_VALID = 1u << 31;
constexpr uint32_t SW_UMETA_BIG = 1u << 30;
constexpr uint32_t SW_UMETA_RTS = 1u << 29;
constexpr uint32_t SW_UMETA_EPMASK = 0xFFFFu;

// ---------------------------------------------------------------- bulk copy input
struct SwSeg {        // one contiguous piece of a rendezvous/loopback copy
  uint64_t src;
  uint64_t dst;
  uint64_t len;
  uint64_t pad;
};

// Balanced bulk copy: the launch carries its jobs (whole messages, at most SW_BULK_INLINE_JOBS of them)
// in the kernel parameter bank and every CTA takes an equal, contiguous byte range of their
// concatenation -- no segment list in host memory, no idle SMs when the launch has fewer segments
// than the GPU has SMs.  The iterator below is the ONLY place that maps (CTA, progress) to addresses;
// the CPU stand-in backend runs the same code, so every host-logic test exercises it.
constexpr uint32_t SW_BULK_INLINE_JOBS = 96;
struct SwBulkJobArgs {
  uint32_t njobs, stage_bytes, nstages, pad;
  uint64_t share;                        // bytes per CTA (multiple of 1024)
  uint64_t end[SW_BULK_INLINE_JOBS];     // end[j] = total bytes of jobs 0..j
  uint64_t src[SW_BULK_INLINE_JOBS];
  uint64_t dst[SW_BULK_INLINE_JOBS];
};
static_assert(sizeof(SwBulkJobArgs) <= 4096 - 512, "bulk job parameters stay under the classic 4 KiB parameter limit");

struct SwJobRangeIter {
  uint64_t pos, range_end;
  uint32_t j;
  SW_HD void init(const SwBulkJobArgs& a, uint32_t cta) {
    const uint64_t total = a.njobs ? a.end[a.njobs - 1] : 0;
    pos = static_cast<uint64_t>(cta) * a.share;
    if (pos > total) pos = total;
    range_end = pos + a.share;
    if (range_end > total) range_end = total;
    j = 0;
    while (j < a.njobs && a.end[j] <= pos) j++;
  }
  // next piece of at most stage_bytes that stays inside one job and inside this CTA's range
  SW_HD bool next(const SwBulkJobArgs& a, uint64_t& src, uint64_t& dst, uint32_t& bytes) {
    if (pos >= range_end) return false;
    while (a.end[j] <= pos) j++;   // pos < range_end <= end[njobs-1]: terminates inside the array
    const uint64_t begin = j ? a.end[j - 1] : 0;
    const uint64_t stop = a.end[j] < range_end ? a.end[j] : range_end;
    const uint64_t left = stop - pos;
    bytes = left < a.stage_bytes ? static_cast<uint32_t>(left) : a.stage_bytes;
    src = a.src[j] + (pos - begin);
    dst = a.dst[j] + (pos - begin);
    pos += bytes;
    return true;
  }
};

// Match rule of the UCP tag API as used at reference main.cpp:404,1172:
// a receive (tag, mask) accepts a message with sender tag `stag` iff
// ((stag ^ tag) & mask) == 0.
SW_HD static inline bool sw_tag_match(uint64_t stag, uint64_t tag, uint64_t mask) {
  return ((stag ^ tag) & mask) == 0;
}
