/* host/shim_util.c -- coolmic_util_*: meter results -> colours.
 *
 * Same contract as reference src/util.c:59-139: a power in dB or a peak sample becomes a hue on the
 * "default" profile's green -> yellow -> red scale, and alpha/hue/saturation/value becomes a packed
 * 0xAARRGGBB. This is a few bytes of host-side double arithmetic per stream and report, so it stays
 * on the host next to the dB finaliser (cmgpu_finalise): same libm as the reference, hence the same
 * doubles and the same 8-bit truncations; a device sin()/pow() could differ in the last place and
 * flip a colour byte.
 */
#include "shim_internal.h"

#include <math.h>
#include <string.h>

static uint32_t unit_to_byte(double x)
{
    /* clamp to [0, 1], scale, truncate (util.c:30-44) */
    uint32_t v;
    if (x >= 1.)
        x = 1.;
    else if (x <= 0.)
        x = 0.;
    v = (uint32_t)(x * 255.);
    return v > 255u ? 255u : v;
}

coolmic_argb_t coolmic_util_ahsv2argb(double alpha, double hue, double saturation, double value)
{
    /* util.c:59-106 -- note the sector's fractional part is taken of `hue` itself, not of
     * hue / (pi/3), exactly as the reference does */
    const int sector = (int)(double)(hue / (M_PI / 3.));
    const double f = hue - (double)sector;
    const double p = value * (1. - saturation);
    const double q = value * (1. - saturation * f);
    const double t = value * (1. - saturation * (1. - f));
    double r = 0., g = 0., b = 0.;

    switch (sector) {
    case 0: case 6: r = value; g = t;     b = p;     break;
    case 1:         r = q;     g = value; b = p;     break;
    case 2:         r = p;     g = value; b = t;     break;
    case 3:         r = p;     g = q;     b = value; break;
    case 4:         r = t;     g = p;     b = value; break;
    case 5:         r = value; g = p;     b = q;     break;
    default:        break;      /* outside [0, 2 pi]: black, like the reference */
    }
    return (unit_to_byte(alpha) << 24) + (unit_to_byte(r) << 16) + (unit_to_byte(g) << 8) + unit_to_byte(b);
}

double coolmic_util_power2hue(double power, const char *profile)
{
    /* util.c:110-122: green below -20 dB, red at 0 dB, sin^2 ramp in between */
    if (!profile || strcmp(profile, COOLMIC_UTIL_PROFILE_DEFAULT) != 0)
        return 0.;
    if (power < -20.)
        return M_PI * 2. / 3.;
    if (power >= 0)
        return 0;
    return pow(sin(M_PI * power / 40.), 2.) * M_PI * 2. / 3.;
}

double coolmic_util_peak2hue(int16_t peak, const char *profile)
{
    /* util.c:126-139: red at full scale, orange above 30000, yellow above 28000, else green */
    if (!profile || strcmp(profile, COOLMIC_UTIL_PROFILE_DEFAULT) != 0)
        return 0.;
    if (peak == -32768 || peak == 32767)
        return 0.;
    if (peak < -30000 || peak > 30000)
        return 0.43;
    if (peak < -28000 || peak > 28000)
        return 1.;
    return M_PI * 2. / 3.;
}
