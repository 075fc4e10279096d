/* oracle/tagmatch.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the tag-matching semantics the reference obtains from
 * libucp (OpenUCX, pinned v1.18.1 in the reference's pyproject.toml:67-68) behind
 *   ucp_tag_send_nbx   reference src/bindings/main.cpp:370, 1136
 *   ucp_tag_recv_nbx   reference src/bindings/main.cpp:404, 1172
 *   ucp_worker_progress reference src/bindings/main.cpp:362, 1127
 * libucp itself is NOT vendored in /root/reference and is absent from this image,
 * so the algorithm below restates the published UCP tag-API contract (SURVEY.md
 * Appendix A) and is pinned against the outcomes the reference's own tests assert
 * (reference tests/test_basic.py, see tests/test_oracle_reference_cases.py).
 * Items no reference test pins (partial masks, truncation, match order among several
 * candidates, per-sender FIFO) follow the documented UCX/MPI rule: "parity unpinned"
 * for those, see DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
 * may use anything in oracle/.
 */
#ifndef ORACLE_TAGMATCH_H
#define ORACLE_TAGMATCH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_ERR_MESSAGE_TRUNCATED = -9, ORC_ERR_CANCELED = -16 };

typedef struct orc_worker orc_worker;

typedef struct orc_match {
  uint64_t op_id;       /* receive that completed */
  uint64_t sender_tag;  /* reference: info->sender_tag (main.cpp:221,1050) */
  uint64_t length;      /* reference: info->length */
  int32_t status;       /* ORC_OK or ORC_ERR_MESSAGE_TRUNCATED */
  uint32_t ep;          /* endpoint the message came from (not visible to the reference API) */
  uint64_t user;        /* opaque value supplied with the message (e.g. a rendezvous id) */
} orc_match;

orc_worker* orc_worker_new(void);
void orc_worker_free(orc_worker* w);

/* ((stag ^ tag) & mask) == 0 */
int orc_tag_match(uint64_t stag, uint64_t tag, uint64_t mask);

/* Post a receive (reference main.cpp:393-425 / 1159-1194).  If an unexpected message
 * matches, the EARLIEST-ARRIVED one is consumed: its bytes are copied into buf
 * (unless truncated), *out is filled and 1 is returned.  Otherwise the receive is
 * appended to the posted queue and 0 is returned. */
int orc_post_recv(orc_worker* w, uint64_t op_id, uint64_t tag, uint64_t mask, void* buf, uint64_t cap,
                  orc_match* out);

/* A message arrives on endpoint ep.  If posted receives match, the EARLIEST-POSTED
 * one completes (copy, *out filled, returns 1).  Otherwise the message (a private
 * copy of its bytes) is appended to the unexpected queue and 0 is returned.
 * data may be NULL with len > 0 for "payload held elsewhere" (rendezvous): then no
 * bytes are copied and `user` identifies the payload. */
int orc_arrive(orc_worker* w, uint32_t ep, uint64_t stag, const void* data, uint64_t len, uint64_t user,
               orc_match* out);

/* Close (reference main.cpp:484-507, 680-701): every pending receive fails with
 * "Request canceled".  Writes up to max op ids, returns how many were pending. */
size_t orc_cancel_all(orc_worker* w, uint64_t* op_ids, size_t max);

size_t orc_num_posted(const orc_worker* w);
size_t orc_num_unexpected(const orc_worker* w);

#ifdef __cplusplus
}
#endif
#endif
