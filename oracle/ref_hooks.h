/* Force-included (-include) into the instrumented build of the reference's
 * graphics.cpp only. TEST INFRASTRUCTURE. Declares the hooks the sed-inserted
 * calls refer to; they are defined in ref_driver.cpp. */
#pragma once
struct shader_struct_a2v;
extern "C" {
void href_hook_a2v(int face, int nth, const shader_struct_a2v* a);
void href_hook_prim(int face, int fan);
void href_hook_bbox(void);
void href_hook_inside(void);
void href_hook_frag(int x, int y);
}
