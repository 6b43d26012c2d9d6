/* Minimal stand-in so /root/reference/api/inc/cactus_params_parser.h parses without libxml2 (not installed
 * here). Only pointer typedefs are needed; no libxml2 function is ever called by the checkers. */
#ifndef ORACLE_LIBXML_SHIM_H
#define ORACLE_LIBXML_SHIM_H
typedef struct _xmlDoc *xmlDocPtr;
typedef struct _xmlNode *xmlNodePtr;
typedef unsigned char xmlChar;
#endif
