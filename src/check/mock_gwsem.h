#ifndef MOCK_GWSEM_H_
#define MOCK_GWSEM_H_
#include "gwsem_b200.h"
struct mock_record {
  int device, src_rows, live, runs, row_filter_rows;
  char path[1024], method[16];
  gwsem_model_spec spec;
  gwsem_job job;
  const int32_t* row_filter;
  void (*check)(void);
};
extern struct mock_record mock;
#endif
