/* stand-in for R_ext/Visibility.h */
#define attribute_visible __attribute__((visibility("default")))
