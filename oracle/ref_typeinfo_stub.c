/* ORACLE-SIDE (test infrastructure): inert objects carrying the names of two typeinfo symbols that the reference's
 * paraformer-online.cpp references through dynamic_cast in ParaformerOnline's CONSTRUCTOR.  oracle/funasr_text_ref_shim.cc
 * never runs that constructor (it needs onnxruntime sessions); the objects only let the shared library load.  The key
 * functions that would emit the real typeinfo live in paraformer.cpp / sensevoice-small.cpp, which need onnxruntime. */
const void* _ZTIN6funasr10ParaformerE[2] = {0, 0};
const void* _ZTIN6funasr15SenseVoiceSmallE[2] = {0, 0};
