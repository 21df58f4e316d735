/* libLLVM.so facade over llvmlite's private LLVM -- test infrastructure (oracle/), see README.md.
 * Exists only so that the UNMODIFIED reference's llvm_ad_rgb variant can start in an image without
 * libLLVM: Dr.Jit dlopens $DRJIT_LIBLLVM_PATH and resolves the LLVM-C API by name
 * (drjit-core/src/llvm_api.cpp:61-170).  Nothing here is reference code, and nothing under
 * mitsuba3_b200/ uses it. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <link.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "shim_tab.inc"

#define EXPORT __attribute__((visibility("default")))
void *shim_slot[SHIM_N];
static void *lite;

/* llvmlite's exported wrappers of the new pass manager (llvmlite/binding/newpassmanagers.py) */
static void *(*PY_CreatePTO)(void);
static void (*PY_DisposePTO)(void *);
static void (*PY_PTOSetLoopVectorization)(void *, int);
static void (*PY_PTOSetSLPVectorization)(void *, int);
static void (*PY_PTOSetLoopUnrolling)(void *, int);
static void *(*PY_CreatePassBuilder)(void *tm, void *pto);
static void (*PY_DisposePassBuilder)(void *);
static void *(*PY_buildPerModuleDefaultPipeline)(void *pb, int speed, int size);
static void (*PY_RunNewModulePassManager)(void *mpm, void *module, void *pb);
static void (*PY_DisposeNewModulePassManger)(void *);

static void die(const char *what) { fprintf(stderr, "llvm shim: %s\n", what); abort(); }

__attribute__((constructor)) static void shim_init(void) {
    const char *path = getenv("B200PT_LLVMLITE_SO");
    if (!path) path = SHIM_LLVMLITE_PATH;
    struct stat st;
    if (stat(path, &st) != 0 || (long) st.st_size != SHIM_LLVMLITE_SIZE) {
        fprintf(stderr, "llvm shim: %s is not the libllvmlite.so this shim was generated for; rebuild oracle/llvm_shim\n", path);
        return;     /* slots stay NULL: Dr.Jit then fails to initialise the LLVM backend instead of crashing later */
    }
    lite = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lite) { fprintf(stderr, "llvm shim: %s\n", dlerror()); return; }
    struct link_map *lm = NULL;
    if (dlinfo(lite, RTLD_DI_LINKMAP, &lm) != 0 || !lm) die("dlinfo failed");
    for (int i = 0; i < SHIM_N; ++i) shim_slot[i] = (void *) (lm->l_addr + shim_off[i]);
#define PY(var, name) *(void **) (&var) = dlsym(lite, name); if (!var) die("missing " name)
    PY(PY_CreatePTO, "LLVMPY_CreatePipelineTuningOptions"); PY(PY_DisposePTO, "LLVMPY_DisposePipelineTuningOptions");
    PY(PY_PTOSetLoopVectorization, "LLVMPY_PTOSetLoopVectorization"); PY(PY_PTOSetSLPVectorization, "LLVMPY_PTOSetSLPVectorization");
    PY(PY_PTOSetLoopUnrolling, "LLVMPY_PTOSetLoopUnrolling");
    PY(PY_CreatePassBuilder, "LLVMPY_CreatePassBuilder"); PY(PY_DisposePassBuilder, "LLVMPY_DisposePassBuilder");
    PY(PY_buildPerModuleDefaultPipeline, "LLVMPY_buildPerModuleDefaultPipeline");
    PY(PY_RunNewModulePassManager, "LLVMPY_RunNewModulePassManager"); PY(PY_DisposeNewModulePassManger, "LLVMPY_DisposeNewModulePassManger");
#undef PY
}

/* Dr.Jit's MCJIT path needs position-independent code and, for lack of a C API, finds the relocation-model field of
 * the TargetMachine by scanning for the three consecutive 32-bit fields {Reloc::Static = 0, CodeModel::Small = 1,
 * CodeGenOpt::Aggressive = 3} and overwriting the first (llvm_mcjit.cpp:52-100).  Since LLVM 17 a 64-bit
 * `LargeDataThreshold` sits between the code model and the optimisation level, so the key no longer matches.
 * The threshold is only consulted for the medium / large code models; setting it to 3 under the small model is
 * inert and restores the {0, 1, 3} pattern exactly at the relocation-model field, which Dr.Jit then sets to PIC. */
EXPORT void *LLVMGetExecutionEngineTargetMachine(void *engine) {
    void *(*real)(void *) = (void *(*)(void *)) shim_slot[SHIM_SLOT_LLVMGetExecutionEngineTargetMachine];
    unsigned *tm = (unsigned *) real(engine);
    if (tm)
        for (int i = 96; i < 192; ++i)
            if (tm[i] <= 1u && tm[i + 1] == 1u && tm[i + 2] == 0u && tm[i + 3] == 0u && tm[i + 4] == 3u) { tm[i + 2] = 3u; break; }
    return tm;
}

/* ---- new pass manager (LLVM-C: llvm-c/Transforms/PassBuilder.h) on llvmlite's wrappers ---------- */
EXPORT void *LLVMCreatePassBuilderOptions(void) { return PY_CreatePTO(); }
EXPORT void LLVMDisposePassBuilderOptions(void *o) { PY_DisposePTO(o); }
EXPORT void LLVMPassBuilderOptionsSetLoopVectorization(void *o, int v) { PY_PTOSetLoopVectorization(o, v); }
EXPORT void LLVMPassBuilderOptionsSetSLPVectorization(void *o, int v) { PY_PTOSetSLPVectorization(o, v); }
EXPORT void LLVMPassBuilderOptionsSetLoopUnrolling(void *o, int v) { PY_PTOSetLoopUnrolling(o, v); }
/* Dr.Jit only ever asks for "default<O2>" (llvm_core.cpp: DRJIT_RUN_NEW_PASS_MANAGER) */
EXPORT void *LLVMRunPasses(void *module, const char *passes, void *tm, void *opts) {
    int speed = 2;
    if (passes && strstr(passes, "O3")) speed = 3; else if (passes && strstr(passes, "O1")) speed = 1; else if (passes && strstr(passes, "O0")) speed = 0;
    void *pb = PY_CreatePassBuilder(tm, opts);
    void *mpm = PY_buildPerModuleDefaultPipeline(pb, speed, 0);
    PY_RunNewModulePassManager(mpm, module, pb);
    PY_DisposeNewModulePassManger(mpm);
    PY_DisposePassBuilder(pb);
    return NULL;     /* LLVMErrorRef: success */
}

/* ---- pieces llvmlite's LLVM was built without: the disassembler (only used for Dr.Jit's trace-level
 *      assembly dumps) and the legacy LICM pass (unused once the new pass manager is available) ---- */
EXPORT void LLVMInitializeX86Disassembler(void) {}
EXPORT void *LLVMCreateDisasm(const char *t, void *d, int tt, void *a, void *b) { (void) t; (void) d; (void) tt; (void) a; (void) b; return NULL; }
EXPORT void LLVMDisasmDispose(void *d) { (void) d; }
EXPORT int LLVMSetDisasmOptions(void *d, unsigned long o) { (void) d; (void) o; return 0; }
EXPORT size_t LLVMDisasmInstruction(void *d, unsigned char *b, unsigned long n, unsigned long pc, char *out, size_t sz) {
    (void) d; (void) b; (void) n; (void) pc; if (sz) out[0] = 0; return 0;
}
EXPORT void LLVMAddLICMPass(void *pm) { (void) pm; }
