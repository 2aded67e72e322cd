// TEST INFRASTRUCTURE (oracle/): stand-in for SoapySDR/Formats.hpp.
#ifndef AERODDC_SOAPY_FORMATS_HPP
#define AERODDC_SOAPY_FORMATS_HPP
#define SOAPY_SDR_CF32 "CF32"
#define SOAPY_SDR_CS16 "CS16"
#define SOAPY_SDR_CU8 "CU8"
#endif
