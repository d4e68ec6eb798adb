// STAND-IN for cnr_class_loader (JRL-CARI-CNR-UNIBS/cnr_class_loader, absent from this image): the registration
// macro records nothing.  Test infrastructure for oracle/_ref only.
#pragma once
#define CLASS_LOADER_REGISTER_CLASS(Derived, Base)
