/* TEST INFRASTRUCTURE ONLY (oracle/): declarations of the libzmq entry points and constants
 * that /root/reference/publish/zmqpublisher.cpp:14-73 uses. The harness (ref_harness.cpp)
 * defines them as an in-memory sink that keeps each 3-frame message. */
#ifndef AERODDC_ORACLE_ZMQ_SHIM_H
#define AERODDC_ORACLE_ZMQ_SHIM_H
#include <cstddef>
#define ZMQ_PUB 1
#define ZMQ_SNDMORE 2
#define ZMQ_RECONNECT_IVL 18
#define ZMQ_RECONNECT_IVL_MAX 21
#define ZMQ_TCP_KEEPALIVE 34
#define ZMQ_TCP_KEEPALIVE_CNT 35
#define ZMQ_TCP_KEEPALIVE_IDLE 36
#define ZMQ_TCP_KEEPALIVE_INTVL 37
extern "C" {
void *zmq_ctx_new(void);
void *zmq_socket(void *ctx, int type);
int zmq_setsockopt(void *s, int option, const void *optval, size_t optvallen);
int zmq_bind(void *s, const char *addr);
int zmq_connect(void *s, const char *addr);
int zmq_send(void *s, const void *buf, size_t len, int flags);
}
#endif
