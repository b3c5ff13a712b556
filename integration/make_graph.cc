// make_graph.cc -- TEST INFRASTRUCTURE: writes a small synthetic decoding graph for the drop-in
// test of the unchanged reference decoder (the reference bundles no HCLG).
//
//   make_graph <dir> <num_pdfs> [n_words] [seed] [first_word_label]
//
// first_word_label (default 1): the graph's words are labels first .. first + n_words - 1 of a symbol
// table the caller supplies instead of the words.txt written here (the delta-LM test uses the
// reference's test/data/lm.words.txt so that its G.pfst / lm.1order.bin fixtures apply).
//
// writes <dir>/HCLG.fst      fst::ConstFst<StdArc>: a word loop; every word is a left-to-right chain
//                            of three emitting states with self-loops, input labels are
//                            transition-ids (>= 1), the word label sits on the entry arc
//        <dir>/words.txt     symbol table ("<eps> 0", words, "<s>", "</s>")   src/symbol_table.cc:17-52
//        <dir>/tid2pdf.bin   VEC0 of int32: transition-id -> pdf-id            src/am.cc:56-61
// Built against the OpenFst headers vendored in the reference tree (nothing is copied).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include <fst/const-fst.h>
#include <fst/vector-fst.h>

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <dir> <num_pdfs> [n_words] [seed] [first_word_label]\n", argv[0]);
    return 2;
  }
  const std::string dir = argv[1];
  const int num_pdfs = atoi(argv[2]);
  const int n_words = argc > 3 ? atoi(argv[3]) : 12;
  uint64_t rng = argc > 4 ? strtoull(argv[4], nullptr, 10) : 20261018ull;
  const int first_label = argc > 5 ? atoi(argv[5]) : 1;
  auto next = [&rng]() {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng >> 33);
  };

  typedef fst::StdArc Arc;
  fst::VectorFst<Arc> g;
  const int loop = g.AddState();
  g.SetStart(loop);
  g.SetFinal(loop, Arc::Weight::One());
  std::vector<int32_t> tid2pdf(1, 0);                     // transition-id 0 is epsilon
  auto new_tid = [&]() {
    tid2pdf.push_back((int32_t)(next() % (uint32_t)num_pdfs));
    return (int)tid2pdf.size() - 1;
  };
  // costs chosen so that the acoustic scores (scaled by 0.1 in the decoder) decide the path: staying
  // in a state is expensive, so an utterance walks through many words
  const float word_cost = 0.1f, loop_cost = 1.5f;
  for (int w = 0; w < n_words; ++w) {
    const int word_label = first_label + w;
    int prev = loop;
    for (int s = 0; s < 3; ++s) {
      const int st = g.AddState();
      // forward arc into the state (the first one carries the word and its LM cost)
      g.AddArc(prev, Arc(new_tid(), s == 0 ? word_label : 0, s == 0 ? word_cost : 0.0f, st));
      g.AddArc(st, Arc(new_tid(), 0, loop_cost, st));       // self-loop
      prev = st;
    }
    g.AddArc(prev, Arc(new_tid(), 0, 0.0f, loop));          // leave the word
  }
  fst::ConstFst<Arc> cg(g);
  if (!cg.Write(dir + "/HCLG.fst")) return 1;

  FILE *f = fopen((dir + "/words.txt").c_str(), "w");
  if (!f) return 1;
  fprintf(f, "<eps> 0\n");
  for (int w = 0; w < n_words; ++w) fprintf(f, "word%02d %d\n", w, w + 1);
  fprintf(f, "<s> %d\n</s> %d\n", n_words + 1, n_words + 2);
  fclose(f);

  f = fopen((dir + "/tid2pdf.bin").c_str(), "wb");
  if (!f) return 1;
  const int32_t dim = (int32_t)tid2pdf.size(), bytes = 4 * dim + 4;
  fwrite("VEC0", 1, 4, f);
  fwrite(&bytes, 4, 1, f);
  fwrite(&dim, 4, 1, f);
  fwrite(tid2pdf.data(), 4, tid2pdf.size(), f);
  fclose(f);
  printf("states %d transition-ids %d words %d\n", (int)g.NumStates(), dim - 1, n_words);
  return 0;
}
