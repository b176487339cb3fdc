// Geometry conventions of the VO front end (documentation contract; mirrors the role of the
// reference's src/geometry_conventions.h:11-28).
//
//  * Pixels: OpenCV 0-based, homogeneous x = [u, v, 1]^T.
//  * Two views: x1 in the previous frame, x2 in the current frame.
//      fundamental   x2^T F x1 = 0   (pixel coordinates)
//      essential     x^2^T E x^1 = 0 with x^ = K^-1 x,  E = [t]_x R,  F = K^-T E K^-1.
//  * Relative motion (OpenCV recoverPose convention):  x_c2 = R x_c1 + t,  |t| = 1
//      i.e. (R, t) maps camera-1 coordinates into camera-2 coordinates.
//  * Stored pose: camera -> world, T_wc = [R_wc | t_wc],  x_w = R_wc x_c + t_wc,  C_w = t_wc.
//      World -> camera: R_cw = R_wc^T, t_cw = -R_wc^T t_wc.
//  * Chaining: T_w_c2 = T_w_c1 * T_c1_c2 with T_c1_c2 = (R^T, -R^T t) and the translation scaled
//      by a fixed visualisation scale (monocular scale is unobservable).
#pragma once
