// spinn_host.h -- host-side helpers shared by the translation units of libspinn_b200.so
#pragma once
int spinn_fail_msg(int code, const char* msg);   // records the thread-local error text, returns code
int spinn_sm_count();
