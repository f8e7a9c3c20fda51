"""B200-native (sm_100a) inference hot path of UpliftingTableTennis: frame stack -> ball-detection
heatmap network -> heatmap peak / sub-pixel decode -> 2D->3D uplifting transformer, behind the
reference's hub / interface API.  Kernels live in csrc/ and are reached through the C ABI of
include/ttk.h (ctypes); there is no CPU fallback."""
__version__ = '0.1.0'
