/* oracle/shim_host/MetalKit/MetalKit.h -- TEST INFRASTRUCTURE. Empty stand-in: the reference's host headers import it
 * (RT_Metal/Metal/Common.hh:20-22) but the BVH builder uses nothing from it. */
