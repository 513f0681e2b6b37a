/* Build-only shim so the reference's gotoh.cpp (whose non-Python branch expects
 * Ruby 1.8's ruby.h) can be compiled UNMODIFIED as the parity oracle.
 * TEST INFRASTRUCTURE ONLY - declares just the handful of symbols that
 * /root/reference/micall/alignment/gotoh.cpp:743-797 names. */
#ifndef GOTOH_ORACLE_STUB_RUBY_H
#define GOTOH_ORACLE_STUB_RUBY_H
typedef unsigned long VALUE;
struct RString { char* ptr; };
#define RSTRING(s) ((RString*)(s))
#define RSTRING_PTR(s) (RSTRING(s)->ptr)
#define NUM2INT(x) ((int)(x))
extern "C" {
VALUE rb_ary_new3(long n, ...);
VALUE rb_str_new2(const char* s);
void rb_define_global_function(const char* name, VALUE (*fn)(...), int argc);
}
#endif
