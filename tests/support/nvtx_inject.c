/* Test-only NVTX injection library: NVTX3 (header-only, inside libdtc_b200.so) loads the library named by NVTX_INJECTION64_PATH on
 * its first call and asks it for callbacks.  This one appends the name of every pushed range to the file named by DTC_NVTX_LOG, so
 * tests/test_abi.py can see that the C-ABI entry points open their ranges without a profiler in the image. */
#include <stdio.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>

static int on_push(const char* name) {
  const char* path = getenv("DTC_NVTX_LOG");
  if (path) {
    FILE* f = fopen(path, "a");
    if (f) { fprintf(f, "push %s\n", name); fclose(f); }
  }
  return 0;
}
static int on_pop(void) {
  const char* path = getenv("DTC_NVTX_LOG");
  if (path) {
    FILE* f = fopen(path, "a");
    if (f) { fprintf(f, "pop\n"); fclose(f); }
  }
  return 0;
}
int InitializeInjectionNvtx2(NvtxGetExportTableFunc_t get_export_table) {
  const NvtxExportTableCallbacks* cb = (const NvtxExportTableCallbacks*)get_export_table(NVTX_ETID_CALLBACKS);
  NvtxFunctionTable table = 0;
  unsigned int size = 0;
  if (!cb || !cb->GetModuleFunctionTable(NVTX_CB_MODULE_CORE, &table, &size)) return 0;
  if (size <= NVTX_CBID_CORE_RangePop) return 0;
  *table[NVTX_CBID_CORE_RangePushA] = (NvtxFunctionPointer)on_push;
  *table[NVTX_CBID_CORE_RangePop] = (NvtxFunctionPointer)on_pop;
  return 1;
}
