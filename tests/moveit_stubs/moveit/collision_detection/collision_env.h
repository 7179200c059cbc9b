#include <moveit_stub_all.h>  // tests/moveit_stubs: declarations-only stand-in, see moveit_stub_all.h
