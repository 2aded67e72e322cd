/* TEST INFRASTRUCTURE (oracle/): the slice of libcorrect's C API that decode/jconvolutionalcodec.cpp calls.
 * libcorrect (quiet/libcorrect, pinned f5a28c74... in the reference README) is not vendored in /root/reference; its
 * convolutional decoder is restated in oracle/viterbi_restated.c (PARITY UNPINNED - see that file's header). */
#ifndef AERODDC_CORRECT_H
#define AERODDC_CORRECT_H
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>
typedef struct correct_convolutional correct_convolutional;
typedef uint16_t correct_convolutional_polynomial_t;
typedef uint8_t correct_convolutional_soft_t;
correct_convolutional* correct_convolutional_create(size_t inv_rate, size_t order, const correct_convolutional_polynomial_t* poly);
void correct_convolutional_destroy(correct_convolutional* conv);
size_t correct_convolutional_encode_len(correct_convolutional* conv, size_t msg_len);
size_t correct_convolutional_encode(correct_convolutional* conv, const uint8_t* msg, size_t msg_len, uint8_t* encoded);
ssize_t correct_convolutional_decode(correct_convolutional* conv, const uint8_t* encoded, size_t num_encoded_bits, uint8_t* msg);
ssize_t correct_convolutional_decode_soft(correct_convolutional* conv, const correct_convolutional_soft_t* encoded, size_t num_encoded_bits, uint8_t* msg);
#endif
