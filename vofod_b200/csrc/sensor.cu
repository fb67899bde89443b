// Real-sensor front end and the host-format packer of N1 — host-only helpers of the C ABI.
//   vofod_mask_mangle    VoFOD::load_mask (vofod_nodelet.cpp:506-560) on a decoded image (PNG decoding stays with the caller's OpenCV)
//   vofod_make_xyz_lut   VoFOD::initialize_sensor_lut (:358-371) = ouster::make_xyz_lut + float cast + per-column normalisation
//   vofod_sim_xyz_lut    VoFOD::initialize_sensor_lut_simulation (:374-420)
//   vofod_pack_ouster    48-byte ouster_ros::Point records (include/vofod/types.h:7, ouster_ros/point.h) -> packed 20-byte vofod_pt
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/vofod_cuda.h"

extern "C" {

/* load_mask (:506-560).  img: rows x cols u8, row-major, or NULL when the file is missing / unreadable.
 *  - wrong dimensions or NULL: every pixel valid (the final ret.resize(exp_cols*exp_rows, 1), :558);
 *  - mangle == 0: mask copied as it is (:529-534);
 *  - mangle != 0 (forced on in simulation, :196): out[((v + pixel_shift_by_row[u]) % W) * H + u] = img[u * W + v]  (:537-548) — the
 *    staggered, column-major order in which the Ouster driver emits the points of one scan.
 * A shift that makes (v + shift) negative indexes out of range in the reference (std::vector::at throws): VOFOD_E_INVALID here. */
int vofod_mask_mangle(const uint8_t* img, int cols, int rows, int W, int H, int mangle, const int32_t* pixel_shift_by_row, uint8_t* out)
{
  if (!out || W <= 0 || H <= 0)
    return VOFOD_E_INVALID;
  const size_t n = (size_t)W * H;
  if (!img || cols != W || rows != H)
  {
    memset(out, 1, n);
    return VOFOD_OK;
  }
  if (!mangle)
  {
    memcpy(out, img, n);
    return VOFOD_OK;
  }
  if (!pixel_shift_by_row)
    return VOFOD_E_INVALID;
  memset(out, 0, n);  // ret.resize(cols*rows): value-initialised, then every index is written exactly once
  for (int u = 0; u < H; u++)
    for (int v = 0; v < W; v++)
    {
      const int t = (v + pixel_shift_by_row[u]) % W;  // int arithmetic, then converted to size_t (:543)
      if (t < 0)
        return VOFOD_E_INVALID;
      out[(size_t)t * H + u] = img[(size_t)u * W + v];
    }
  return VOFOD_OK;
}

/* initialize_sensor_lut (:358-371): ouster::make_xyz_lut (ouster_example 2.x lidar_scan.cpp — third-party, not in the reference tree:
 * restated from its published formula) in double, cast to float, directions normalised per column in fp32.
 *   encoder(u,v) = 2 pi - v * 2 pi / w;  azimuth = -azimuth_deg[u] pi/180;  altitude = altitude_deg[u] pi/180
 *   dir = (cos(enc+az) cos(alt), sin(enc+az) cos(alt), sin(alt));  off = (cos(enc) - dir.x, sin(enc) - dir.y, -dir.z) * beam_origin_mm
 *   dir = R dir;  off = R off + t   (lidar_to_sensor_transform, row-major 4x4);  both scaled by range_unit
 * Ray id = u * w + v (row u, column v), as the nodelet indexes the LUT (:1446).  dirs / offs: 3 floats per ray. */
int vofod_make_xyz_lut(int w, int h, double range_unit, double lidar_origin_to_beam_origin_mm, const double* transform4x4, const double* azimuth_angles_deg,
                       const double* altitude_angles_deg, float* dirs, float* offs)
{
  if (w <= 0 || h <= 0 || !azimuth_angles_deg || !altitude_angles_deg || !dirs || !offs)
    return VOFOD_E_INVALID;
  static const double ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const double* T = transform4x4 ? transform4x4 : ident;
  const double azimuth_radians = M_PI * 2.0 / w;
  for (int v = 0; v < w; v++)
    for (int u = 0; u < h; u++)
    {
      const size_t i = (size_t)u * w + v;
      const double enc = 2.0 * M_PI - (v * azimuth_radians);
      const double az = -azimuth_angles_deg[u] * M_PI / 180.0;
      const double alt = altitude_angles_deg[u] * M_PI / 180.0;
      const double d[3] = {cos(enc + az) * cos(alt), sin(enc + az) * cos(alt), sin(alt)};
      const double o[3] = {(cos(enc) - d[0]) * lidar_origin_to_beam_origin_mm, (sin(enc) - d[1]) * lidar_origin_to_beam_origin_mm, (-d[2]) * lidar_origin_to_beam_origin_mm};
      double dr[3], orr[3];
      for (int a = 0; a < 3; a++)
      {
        dr[a] = (T[4 * a] * d[0] + T[4 * a + 1] * d[1] + T[4 * a + 2] * d[2]) * range_unit;
        orr[a] = (T[4 * a] * o[0] + T[4 * a + 1] * o[1] + T[4 * a + 2] * o[2] + T[4 * a + 3]) * range_unit;
      }
      // {direction.cast<float>(), offset.cast<float>()}; directions.colwise().normalize()  (:368-369): v /= sqrt(x^2 + y^2 + z^2) in fp32
      const float fx = (float)dr[0], fy = (float)dr[1], fz = (float)dr[2];
      volatile float n2 = fx * fx;
      n2 = n2 + fy * fy;
      n2 = n2 + fz * fz;
      const float nrm = sqrtf(n2);
      dirs[3 * i] = fx / nrm;
      dirs[3 * i + 1] = fy / nrm;
      dirs[3 * i + 2] = fz / nrm;
      offs[3 * i] = (float)orr[0];
      offs[3 * i + 1] = (float)orr[1];
      offs[3 * i + 2] = (float)orr[2];
    }
  return VOFOD_OK;
}

/* initialize_sensor_lut_simulation (:374-420): yaw = col * 2 pi / (W - 1), pitch = -vfov/2 + row * vfov / (H - 1) with vfov the FLOAT
 * member m_sensor_vfov widened to double; double trigonometry stored as float, NOT renormalised; offsets zero; ray id = col + row * W */
int vofod_sim_xyz_lut(int w, int h, float vfov, float* dirs, float* offs)
{
  if (w < 2 || h < 2 || !dirs)
    return VOFOD_E_INVALID;
  const double minAngle = 0.0, maxAngle = 2.0 * M_PI;
  const double verticalMinAngle = -vfov / 2.0, verticalMaxAngle = vfov / 2.0;
  const double yAngle_step = (maxAngle - minAngle) / (w - 1), pAngle_step = (verticalMaxAngle - verticalMinAngle) / (h - 1);
  for (int row = 0; row < h; row++)
    for (int col = 0; col < w; col++)
    {
      const double yAngle = col * yAngle_step + minAngle, pAngle = row * pAngle_step + verticalMinAngle;
      float* d = dirs + 3 * ((size_t)col + (size_t)row * w);
      d[0] = (float)(cos(pAngle) * cos(yAngle));
      d[1] = (float)(cos(pAngle) * sin(yAngle));
      d[2] = (float)sin(pAngle);
    }
  if (offs)
    memset(offs, 0, (size_t)w * h * 12);
  return VOFOD_OK;
}

/* ouster_ros::Point (ouster_ros/point.h, the point type of the nodelet's input cloud, include/vofod/types.h:7): EIGEN_ALIGN16
 * { float x, y, z, <pad>; float intensity; uint32_t t; uint16_t reflectivity; uint8_t ring; uint16_t ambient; uint32_t range; } =
 * 48 bytes, field offsets 0, 4, 8, 16, 20, 24, 26, 28, 32.  `stride` / the two offsets let other builds of the driver be packed too
 * (0 = the layout above). */
int vofod_pack_ouster(const void* points, size_t n, size_t stride, size_t intensity_offset, size_t range_offset, vofod_pt* out)
{
  if ((n && !points) || !out)
    return VOFOD_E_INVALID;
  if (stride == 0)
  {
    stride = 48;
    intensity_offset = 16;
    range_offset = 32;
  }
  if (intensity_offset + 4 > stride || range_offset + 4 > stride || stride < 16)
    return VOFOD_E_INVALID;
  const unsigned char* b = (const unsigned char*)points;
  for (size_t i = 0; i < n; i++, b += stride)
  {
    memcpy(&out[i].x, b, 12);
    memcpy(&out[i].intensity, b + intensity_offset, 4);
    memcpy(&out[i].range_mm, b + range_offset, 4);
  }
  return VOFOD_OK;
}
}  // extern "C"
