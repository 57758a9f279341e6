/* Minimal C host for the vilbert_b200 C ABI (include/vilbert_b200.h): what a non-Python replacement of the reference's queue
 * worker (callback/main, worker.py:542-673) links against.  It needs no GPU to run: vb200_create audits the configuration and the
 * checkpoint BEFORE touching a device, so a wrong config / an incomplete state_dict comes back as a status + message.
 *
 *   gcc -std=c11 -Iinclude examples/host_min.c -o host_min -ldl && ./host_min vilbert-multi-task_b200/libvilbert_b200.so
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "vilbert_b200.h"

typedef int (*abi_fn)(void);
typedef int (*create_fn)(const char*, int64_t, const vb200_tensor*, const vb200_options*, vb200_handle*);
typedef const char* (*err_fn)(vb200_handle);
typedef int (*destroy_fn)(vb200_handle);

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s path/to/libvilbert_b200.so\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    abi_fn abi = (abi_fn)dlsym(lib, "vb200_abi_version");
    create_fn create = (create_fn)dlsym(lib, "vb200_create");
    err_fn last_error = (err_fn)dlsym(lib, "vb200_last_error");
    destroy_fn destroy = (destroy_fn)dlsym(lib, "vb200_destroy");
    if (!abi || !create || !last_error || !destroy) { fprintf(stderr, "missing symbol\n"); return 2; }
    printf("abi %d (header %d)\n", abi(), VB200_ABI_VERSION);
    if (abi() != VB200_ABI_VERSION) return 1;

    vb200_handle h = NULL;
    float dummy[4] = {0};
    vb200_tensor t = {"bert.embeddings.LayerNorm.weight", VB200_F32, 1, {4, 0}, dummy};
    int rc = create("{ this is not json", 1, &t, NULL, &h);
    printf("bad config       -> %d : %s\n", rc, last_error(NULL));
    if (rc != VB200_ERR_CONFIG || h != NULL) return 1;

    /* a syntactically valid ViLBERT config with an empty state_dict: the audit names the first missing tensor */
    const char* cfg =
        "{\"hidden_size\":768,\"num_hidden_layers\":12,\"num_attention_heads\":12,\"intermediate_size\":3072,"
        "\"hidden_act\":\"gelu\",\"max_position_embeddings\":512,\"type_vocab_size\":2,\"vocab_size\":30522,"
        "\"v_feature_size\":2048,\"v_target_size\":1601,\"v_hidden_size\":1024,\"v_num_hidden_layers\":6,"
        "\"v_num_attention_heads\":8,\"v_intermediate_size\":1024,\"bi_hidden_size\":1024,\"bi_num_attention_heads\":8,"
        "\"bi_intermediate_size\":1024,\"bi_attention_type\":1,\"v_biattention_id\":[0,1,2,3,4,5],"
        "\"t_biattention_id\":[6,7,8,9,10,11],\"task_specific_tokens\":true}";
    rc = create(cfg, 1, &t, NULL, &h);
    printf("empty checkpoint -> %d : %s\n", rc, last_error(NULL));
    if (rc != VB200_ERR_CHECKPOINT || h != NULL) return 1;
    if (strstr(last_error(NULL), "missing key") == NULL) return 1;   /* the message names the upstream key */
    dlclose(lib);
    printf("ok\n");
    return 0;
}
