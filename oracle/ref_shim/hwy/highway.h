// shim (oracle/_ref build only): Highway is absent from this image; the reference headers compiled here only name it
// inside `#if HWY_HAVE_CONSTEXPR_LANES` blocks.
#pragma once
#define HWY_HAVE_CONSTEXPR_LANES 0
