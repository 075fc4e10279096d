/* oracle/tagmatch.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See tagmatch.h.
 *
 * Plain doubly-linked FIFO lists, linear search: the simplest possible statement of
 * the UCP tag-matching contract:
 *   1. matching domain = one worker; no source field;
 *   2. recv (tag, mask) matches stag iff ((stag ^ tag) & mask) == 0;
 *   3. arriving message -> earliest-posted matching recv;
 *      new recv -> earliest-arrived matching unexpected message;
 *      per-endpoint FIFO is the caller's job (call orc_arrive in send order);
 *   4. len <= cap: success, `length` = len, bytes beyond untouched;
 *      len > cap: "Message truncated", message consumed, buffer untouched;
 *   7. close: every pending receive is cancelled.
 */
#include "tagmatch.h"

#include <stdlib.h>
#include <string.h>

typedef struct orc_recv {
  struct orc_recv *prev, *next;
  uint64_t op_id, tag, mask, cap;
  void* buf;
} orc_recv;

typedef struct orc_msg {
  struct orc_msg *prev, *next;
  uint64_t stag, len, user;
  uint32_t ep;
  void* data; /* private copy, or NULL for payload-held-elsewhere */
} orc_msg;

struct orc_worker {
  orc_recv *p_head, *p_tail; /* posted receives, post order */
  orc_msg *u_head, *u_tail;  /* unexpected messages, arrival order */
  size_t n_posted, n_unexp;
};

orc_worker* orc_worker_new(void) { return (orc_worker*)calloc(1, sizeof(orc_worker)); }

void orc_worker_free(orc_worker* w) {
  if (!w) return;
  for (orc_recv* r = w->p_head; r;) {
    orc_recv* n = r->next;
    free(r);
    r = n;
  }
  for (orc_msg* m = w->u_head; m;) {
    orc_msg* n = m->next;
    free(m->data);
    free(m);
    m = n;
  }
  free(w);
}

int orc_tag_match(uint64_t stag, uint64_t tag, uint64_t mask) { return ((stag ^ tag) & mask) == 0; }

static void fill(orc_match* out, uint64_t op_id, uint64_t stag, uint64_t len, uint64_t cap, uint32_t ep,
                 uint64_t user, void* buf, const void* data) {
  out->op_id = op_id;
  out->sender_tag = stag;
  out->length = len;
  out->ep = ep;
  out->user = user;
  if (len > cap) {
    out->status = ORC_ERR_MESSAGE_TRUNCATED;
  } else {
    out->status = ORC_OK;
    if (data && len) memcpy(buf, data, (size_t)len);
  }
}

int orc_post_recv(orc_worker* w, uint64_t op_id, uint64_t tag, uint64_t mask, void* buf, uint64_t cap,
                  orc_match* out) {
  for (orc_msg* m = w->u_head; m; m = m->next) {
    if (!orc_tag_match(m->stag, tag, mask)) continue;
    fill(out, op_id, m->stag, m->len, cap, m->ep, m->user, buf, m->data);
    if (m->prev) m->prev->next = m->next; else w->u_head = m->next;
    if (m->next) m->next->prev = m->prev; else w->u_tail = m->prev;
    free(m->data);
    free(m);
    w->n_unexp--;
    return 1;
  }
  orc_recv* r = (orc_recv*)calloc(1, sizeof(orc_recv));
  r->op_id = op_id;
  r->tag = tag;
  r->mask = mask;
  r->buf = buf;
  r->cap = cap;
  r->prev = w->p_tail;
  if (w->p_tail) w->p_tail->next = r; else w->p_head = r;
  w->p_tail = r;
  w->n_posted++;
  return 0;
}

int orc_arrive(orc_worker* w, uint32_t ep, uint64_t stag, const void* data, uint64_t len, uint64_t user,
               orc_match* out) {
  for (orc_recv* r = w->p_head; r; r = r->next) {
    if (!orc_tag_match(stag, r->tag, r->mask)) continue;
    fill(out, r->op_id, stag, len, r->cap, ep, user, r->buf, data);
    if (r->prev) r->prev->next = r->next; else w->p_head = r->next;
    if (r->next) r->next->prev = r->prev; else w->p_tail = r->prev;
    free(r);
    w->n_posted--;
    return 1;
  }
  orc_msg* m = (orc_msg*)calloc(1, sizeof(orc_msg));
  m->stag = stag;
  m->len = len;
  m->user = user;
  m->ep = ep;
  if (data && len) {
    m->data = malloc((size_t)len);
    memcpy(m->data, data, (size_t)len);
  }
  m->prev = w->u_tail;
  if (w->u_tail) w->u_tail->next = m; else w->u_head = m;
  w->u_tail = m;
  w->n_unexp++;
  return 0;
}

size_t orc_cancel_all(orc_worker* w, uint64_t* op_ids, size_t max) {
  size_t n = 0;
  for (orc_recv* r = w->p_head; r;) {
    orc_recv* nx = r->next;
    if (n < max && op_ids) op_ids[n] = r->op_id;
    n++;
    free(r);
    r = nx;
  }
  w->p_head = w->p_tail = NULL;
  w->n_posted = 0;
  return n;
}

size_t orc_num_posted(const orc_worker* w) { return w->n_posted; }
size_t orc_num_unexpected(const orc_worker* w) { return w->n_unexp; }
