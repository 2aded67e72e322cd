/* TEST INFRASTRUCTURE (oracle/): stand-in for SoapySDR/Constants.h (publisher.cpp:33-38,254). */
#ifndef AERODDC_SOAPY_CONSTANTS_H
#define AERODDC_SOAPY_CONSTANTS_H
#define SOAPY_SDR_TX 0
#define SOAPY_SDR_RX 1
#endif
