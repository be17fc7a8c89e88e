// Minimal stand-in for <boost/predef.h>: only the OS-detection macro a uBLAS-style
// caller of amt::mtm may consult. Used only when real Boost is not installed.
#ifndef B200_COMPAT_BOOST_PREDEF_H
#define B200_COMPAT_BOOST_PREDEF_H
#if defined(__linux__)
#  define BOOST_OS_LINUX_AVAILABLE 1
#  define BOOST_OS_LINUX 1
#endif
#endif
