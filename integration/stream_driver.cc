// stream_driver.cc -- TEST DRIVER for the drop-in C API (src/ce_stt.h): feeds one wav file to
// ce_stt_process in fixed-size pieces, the way src/main.cc:28-52 does with 1024-byte reads, and
// prints the hypothesis text after every call that changed it, then the final one.  Built twice by
// oracle/Makefile: against the reference's own src/ce_stt.cc (stream_ref) and against
// integration/ce_stt_gpu.cc (stream_gpu); tests/test_dropin_decoder.py demands identical output.
//
//   stream_driver <config> <wav> [bytes per call = 1024]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ce_stt.h"

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <config> <wav> [bytes per call]\n", argv[0]);
    return 2;
  }
  const int piece = argc > 3 ? atoi(argv[3]) : 1024;
  ce_stt_t *rec = ce_stt_init(argv[1]);
  if (!rec) {
    fprintf(stderr, "ce_stt_init: %s\n", ce_stt_last_error());
    return 1;
  }
  FILE *fd = fopen(argv[2], "rb");
  ce_wave_format_t fmt;
  if (!fd || !ce_read_pcm_header(fd, &fmt)) {
    fprintf(stderr, "wav: %s\n", fd ? ce_stt_last_error() : "unable to open");
    return 1;
  }
  ce_utt_t *utt = ce_utt_init(rec, &fmt);
  if (!utt) {
    fprintf(stderr, "ce_utt_init: %s\n", ce_stt_last_error());
    return 1;
  }
  std::vector<char> buf(piece);
  std::string last;
  long bytes = 0;
  int calls = 0;
  while (!feof(fd)) {
    const int n = (int)fread(buf.data(), 1, buf.size(), fd);
    if (n == 0) break;
    if (ce_stt_process(utt, buf.data(), n) == CE_STT_FAILED) {
      fprintf(stderr, "ce_stt_process: %s\n", ce_stt_last_error());
      return 1;
    }
    bytes += n;
    ++calls;
    if (last != utt->hyp) {
      last = utt->hyp;
      printf("partial %ld %s\n", bytes, utt->hyp);
    }
  }
  ce_stt_end_of_stream(utt);
  printf("final %d calls %s\n", calls, utt->hyp);
  ce_utt_destroy(utt);
  fclose(fd);
  ce_stt_destroy(rec);
  return 0;
}
