/* Test-infrastructure shim (NOT product code): empty stand-in; the reference includes this header (AudioSDR.h:35-41) but uses nothing from it on the host. */
