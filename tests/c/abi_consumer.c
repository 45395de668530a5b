/* A plain C99 consumer of include/aocr.h: what a cgo / LuaJIT-FFI / JNI binding sees.  Needs no GPU:
 *   - the header compiles as C (no C++ in the signatures),
 *   - the library links and resolves from C,
 *   - the dictionary entry points (src/utils/utils.lua:177-218) run on the host and return the flat child table,
 *   - aocr_create refuses loudly — non-zero status, message through aocr_last_error(NULL) — when the configuration is not
 *     the reference's (dropout != 0) and, on a machine without a device, when there is no GPU (no CPU fallback). */
#include <stdio.h>
#include <string.h>

#include "aocr.h"

int main(void) {
  int32_t* table = NULL;
  int32_t nodes = 0;
  int rc = aocr_trie_from_words("cat\ncar\ndog\n", 0, &table, &nodes);
  if (rc != 0 || table == NULL || nodes < 8) { printf("FAIL trie rc=%d nodes=%d\n", rc, (int)nodes); return 1; }
  printf("trie nodes=%d\n", (int)nodes);
  aocr_trie_free(table);

  aocr_config c;
  memset(&c, 0, sizeof c);
  c.batch_size = 4; c.max_encoder_l = 30; c.max_decoder_l = 10; c.encoder_num_hidden = 512; c.encoder_num_layers = 1;
  c.decoder_num_layers = 2; c.target_vocab_size = 39; c.target_embedding_size = 20; c.input_feed = 1;
  c.dropout = 0.5f; c.learning_rate = 0.1f; c.dp_world = 1;
  aocr_handle* h = NULL;
  rc = aocr_create(&c, 0, &h);
  const char* msg = aocr_last_error(NULL);
  if (rc == 0 || h != NULL || msg == NULL || msg[0] == 0) { printf("FAIL create accepted dropout=0.5\n"); return 1; }
  printf("create refused: rc=%d msg=%s\n", rc, msg);
  printf("OK\n");
  return 0;
}
